"""per-kernel device durations of one fwd+bwd step under torch.profiler (CUPTI), warm caches"""
import sys, torch
sys.path.insert(0, '/root/repo')
import bench, wc_gan_b200 as W
from torch.profiler import profile, ProfilerActivity
dev = torch.device('cuda', 0)
shape = bench.SHAPE
g = int(sys.argv[1]) if len(sys.argv) > 1 else None
xs, dys = bench.make_inputs(torch, shape, 3, dev, 1)
stack = bench.build_stack(W, torch, 'd', 'uconv', shape[-1], dev, g=g)
params = list(stack.parameters())
def step(i):
    x = xs[i % 3].requires_grad_(True)
    y = stack(x)
    torch.autograd.grad(y, [x] + params, dys[i % 3])
for i in range(5): step(i)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(10): step(i)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
agg = {}
for e in evs:
    agg.setdefault(e.name, []).append(e.device_time if hasattr(e, 'device_time') else e.cuda_time)
tot = sum(sum(v) for v in agg.values()) / 10
print(f'total device time per step {tot:.1f} us')
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f'{sum(v)/10:8.1f} us/step  n/step={len(v)/10:4.1f}  avg={sum(v)/len(v):7.1f} us  {k[:90]}')
evs_sorted = sorted(evs, key=lambda e: e.time_range.start)
per = len(evs_sorted) // 10
print('--- one step in launch order (start offset us, duration us)')
t0 = evs_sorted[-per].time_range.start
for e in evs_sorted[-per:]:
    d = e.device_time if hasattr(e, 'device_time') else e.cuda_time
    print(f'{(e.time_range.start - t0):9.1f} {d:7.1f}  {e.name[:100]}')
import time
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(50): step(i)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f'host enqueue {1e6*(t1-t0)/50:.1f} us/step; with sync {1e6*(t2-t0)/50:.1f} us/step')
