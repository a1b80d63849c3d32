"""per-kernel END-to-END timeline of one fwd+bwd step replayed from a CUDA graph (with PDL a kernel's own duration
includes its wait; the gap between consecutive kernel ENDs is what the kernel adds to the step)"""
import sys, torch
sys.path.insert(0, '/root/repo')
import bench, wc_gan_b200 as W
from torch.profiler import profile, ProfilerActivity
dev = torch.device('cuda', 0)
shape = bench.SHAPE
xs, dys = bench.make_inputs(torch, shape, 2, dev, 1)
stack = bench.build_stack(W, torch, 'd', 'uconv', shape[-1], dev)
params = list(stack.parameters())
def step(i):
    x = xs[i % 2].requires_grad_(True)
    y = stack(x)
    return torch.autograd.grad(y, [x] + params, dys[i % 2])
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for i in range(3): step(i)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g, stream=s):
    step(0); step(1)
for _ in range(3): g.replay()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    g.replay()
    torch.cuda.synchronize()
evs = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.end)
half = len(evs) // 2
evs = evs[half:]           # second step of the replay (warm)
prev = evs[0].time_range.start
t0 = prev
print(f'{"end us":>9s} {"adds us":>8s} {"own us":>8s}  kernel')
for e in evs:
    print(f'{e.time_range.end - t0:9.1f} {e.time_range.end - prev:8.1f} {e.time_range.end - e.time_range.start:8.1f}  {e.name[:90]}')
    prev = e.time_range.end
