"""per-kernel END-to-END timeline of one fwd+bwd step replayed from a CUDA graph (with PDL a kernel's own duration
includes its wait; the gap between consecutive kernel ENDs is what the kernel adds to the step)"""
import sys, torch
sys.path.insert(0, '/root/repo')
import bench, wc_gan_b200 as W
from torch.profiler import profile, ProfilerActivity
import os, ctypes
import torch.distributed as dist
from wc_gan_b200._lib import lib as _l
_l.wc_debug_set.argtypes=[ctypes.c_int]
if os.environ.get('WC_DBG'): _l.wc_debug_set(int(os.environ['WC_DBG']))
rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
dev = torch.device('cuda', int(os.environ.get('LOCAL_RANK', 0)))
torch.cuda.set_device(dev)
group = None
if world > 1:
    dist.init_process_group('nccl', device_id=dev); group = dist.group.WORLD
shape = tuple(int(v) for v in os.environ['WC_SHAPE'].split('x')) if os.environ.get('WC_SHAPE') else bench.SHAPE
AN, KK, EE = os.environ.get('WC_AN', 'uconv'), int(os.environ.get('WC_K', 0)) or None, int(os.environ.get('WC_E', 10))
xs, dys = bench.make_inputs(torch, shape, 2, dev, 1 + rank)
stack = bench.build_stack(W, torch, 'd', AN, shape[-1], dev, group=group, K=KK, E=EE)
cls = torch.randint(0, KK, (shape[0], 1), device=dev, dtype=torch.int32) if KK else None
params = list(stack.parameters())
def step(i):
    x = xs[i % 2].requires_grad_(True)
    y = stack(x, cls) if KK else stack(x)
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
if rank == 0: print(f'{"end us":>9s} {"adds us":>8s} {"own us":>8s}  kernel')
for e in evs:
    if rank == 0: print(f'{e.time_range.end - t0:9.1f} {e.time_range.end - prev:8.1f} {e.time_range.end - e.time_range.start:8.1f}  {e.name[:90]}')
    prev = e.time_range.end

if world > 1:
    del g
    torch.cuda.synchronize(); dist.barrier(); os._exit(0)
