"""per-CTA globaltimer marks of the backward chain kernel inside one fwd+bwd step (slots 8..15 of the debug buffer):
8 start | 9 stage-0 jobs done | 10 barrier 1 passed | 11 | 12 | 13 | 14 | 15 end"""
import sys, ctypes, torch
sys.path.insert(0, '/root/repo')
import bench, wc_gan_b200 as W
from wc_gan_b200._lib import lib
lib.wc_debug_set.argtypes = [ctypes.c_int]; lib.wc_debug_marks.argtypes = [ctypes.c_void_p]
dev = torch.device('cuda', 0)
import os
EXTRA = int(os.environ.get('WC_DBG', '0'))
CASES = ((bench.SHAPE, 'uconv', None, 4), ((64, 16, 16, 128), 'ufconv', 10, 4))
if os.environ.get('WC_CFG5'): CASES = (((256, 4, 4, 128), 'ufconv', 200, 15),)
for shape, an, K, EE in CASES:
    xs, dys = bench.make_inputs(torch, shape, 2, dev, 1)
    stack = bench.build_stack(W, torch, 'd', an, shape[-1], dev, K=K, E=EE)
    params = list(stack.parameters())
    cls = torch.randint(0, K, (shape[0], 1), dtype=torch.int32, device=dev) if K else None
    def step(i):
        x = xs[i % 2].requires_grad_(True)
        y = stack(x, cls)
        if os.environ.get('WC_FWD_ONLY'): return y
        return torch.autograd.grad(y, [x] + params, dys[i % 2])
    for i in range(3): step(i)
    torch.cuda.synchronize()
    lib.wc_debug_set(1 | EXTRA)
    for i in range(2): step(i)
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * (160 * 32))()
    lib.wc_debug_marks(buf)
    lib.wc_debug_set(0)
    rows = [[buf[b * 32 + k] for k in range(8, 32)] for b in range(148) if buf[b * 32 + 8]]
    t0 = min(r[0] for r in rows)
    print(f'--- {shape} {an}: {len(rows)} CTAs; per slot: min / median / max us since the first CTA started')
    names = ['start', 'S0 jobs', 'bar1', 'S1 jobs', 'bar2', 'S2 jobs', 'bar3', 'end']
    for k in range(8):
        v = sorted((r[k] - t0) / 1e3 for r in rows if r[k])
        if v: print(f'{names[k]:8s} {v[0]:8.2f} {v[len(v)//2]:8.2f} {v[-1]:8.2f}')
    # inside the last tile each CTA ran (slots 16..27): entry, prologue issued, chunk c landed (c = 0..7), ring drained, epilogue stored
    tn = ['entry', 'issued', 'c0', 'c1', 'c2', 'c3', 'c4', 'c5', 'c6', 'c7', 'drained', 'stored']
    tl = [r for r in rows if r[8] and r[19]]
    # only CTAs whose last tile belongs to the last stage
    print(f'last-stage tiles: {len(tl)} CTAs; SM cycles since tile entry (median)')
    for k in range(12):
        v = sorted((r[8 + k] - r[8]) for r in tl if r[8 + k] >= r[8])
        if v: print(f'  {tn[k]:8s} {v[len(v)//2]:8d}  (max {v[-1]})')
