"""marks inside stats_exchange_kernel under torchrun (slots 28..31: start | record finalized | pushed | all peers arrived)"""
import sys, ctypes, os, torch
sys.path.insert(0, '/root/repo')
import torch.distributed as dist
import bench, wc_gan_b200 as W
from wc_gan_b200._lib import lib
lib.wc_debug_set.argtypes = [ctypes.c_int]; lib.wc_debug_marks.argtypes = [ctypes.c_void_p]
rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
dev = torch.device('cuda', int(os.environ.get('LOCAL_RANK', 0))); torch.cuda.set_device(dev)
dist.init_process_group('nccl', device_id=dev)
xs, dys = bench.make_inputs(torch, bench.SHAPE, 2, dev, 1 + rank)
stack = bench.build_stack(W, torch, 'd', 'uconv', 256, dev, group=dist.group.WORLD)
params = list(stack.parameters())
def step(i):
    x = xs[i % 2].requires_grad_(True)
    y = stack(x)
    return torch.autograd.grad(y, [x] + params, dys[i % 2])
for i in range(5): step(i)
torch.cuda.synchronize(); dist.barrier()
lib.wc_debug_set(1 | int(os.environ.get('WC_DBG', '0')))
for i in range(3): step(i)          # back to back: the steady-state skew between the ranks
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * (160 * 32))()
lib.wc_debug_marks(buf); lib.wc_debug_set(0)
rows = [[buf[b * 32 + k] for k in range(28, 32)] for b in range(128) if buf[b * 32 + 28]]
t0 = min(r[0] for r in rows)
names = ['start', 'finalized', 'pushed', 'peers in']
for r_ in range(world):
    if r_ == rank:
        print(f'rank {rank}: us since first CTA start (min / median / max over {len(rows)} CTAs)')
        for k in range(4):
            v = sorted((r[k] - t0) / 1e3 for r in rows)
            print(f'  {names[k]:10s} {v[0]:7.2f} {v[len(v)//2]:7.2f} {v[-1]:7.2f}', flush=True)
    dist.barrier()
crow = [[buf[b * 32 + k] for k in range(8, 16)] for b in range(148) if buf[b * 32 + 8]]
c0 = min(r[0] for r in crow)
if rank == 0:
    print('rank 0 chain kernel (bwd): start | S0 jobs | exchange+bar done | S1 jobs | bar2 | S2 jobs | bar3 | end  (median us)')
    print('  ' + ' '.join(f'{sorted((r[k] - c0) / 1e3 for r in crow if r[k])[len(crow)//2]:7.2f}' for k in range(8)))
torch.cuda.synchronize(); dist.barrier(); os._exit(0)
