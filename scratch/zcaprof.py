"""kernel durations of one ZCA fwd+bwd step (torch profiler) + Jacobi sweep counts"""
import sys, torch
sys.path.insert(0, '/root/repo')
import bench, wc_gan_b200 as W
from torch.profiler import profile, ProfilerActivity
dev = torch.device('cuda', 0)
for g in (64, 128, None):
    xs, dys = bench.make_inputs(torch, bench.SHAPE, 2, dev, 7)
    st = bench.build_stack(W, torch, 'd', 'uconv', 256, dev, dec='zca', g=g)
    params = list(st.parameters())
    def step(i):
        x = xs[i % 2].requires_grad_(True)
        y = st(x)
        torch.autograd.grad(y, [x] + params, dys[i % 2])
    for i in range(3): step(i)
    torch.cuda.synchronize()
    print('g =', g, 'info (not_pd, bad_class, jacobi sweeps, xchg):', st.info.tolist())
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step(0); torch.cuda.synchronize()
    evs = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
    for e in evs:
        d = e.time_range.end - e.time_range.start
        if d > 8: print(f'   {d:9.1f} us  {e.name[:80]}')
