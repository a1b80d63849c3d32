"""debug: K3 / K6 alone with per-CTA globaltimer marks (start of main loop, end)"""
import sys, ctypes, torch
sys.path.insert(0, '/root/repo')
import wc_gan_b200 as W  # noqa: F401
from wc_gan_b200._lib import lib, ptr
N, HW, C = 256, 256, 256
M = N * HW
dev = torch.device('cuda')
NB = 4
xs = [torch.randn(M, C, device=dev) + 1 for _ in range(NB)]
dys = [torch.randn(M, C, device=dev) for _ in range(NB)]
outs = [torch.empty(M, C, device=dev) for _ in range(2)]
mean = torch.zeros(C, device=dev)
T = torch.randn(C, C, device=dev) / 16; A = torch.randn(C, C, device=dev) / 1e4; b = torch.randn(C, device=dev)
st = torch.cuda.current_stream().cuda_stream
lib.wc_debug_set.argtypes = [ctypes.c_int]; lib.wc_debug_marks.argtypes = [ctypes.c_void_p]
pf = torch.empty(lib.wc_pack_bytes(C, 1, 0) // 4, device=dev); pb = torch.empty(lib.wc_pack_bytes(C, 1, 1) // 4, device=dev)
assert lib.wc_pack_fwd(ptr(T), C, 1, ptr(pf), st) == 0
assert lib.wc_pack_bwd(ptr(T), ptr(A), C, C, 1, ptr(pb), st) == 0
info = torch.zeros(4, dtype=torch.int32, device=dev)
k3 = lambda i: lib.wc_apply_fwd(ptr(xs[i % NB]), N, HW, C, C, ptr(mean), ptr(T), ptr(pf), ptr(b), None, 1, 0, ptr(outs[i % 2]), ptr(info), st)
k6 = lambda i: lib.wc_apply_bwd(ptr(xs[i % NB]), ptr(dys[i % NB]), N, HW, C, C, ptr(mean), ptr(T), ptr(A), ptr(pb), ptr(b), C, None, 1, ptr(outs[i % 2]), st)
def marks(tag):
    buf = (ctypes.c_ulonglong * (160 * 32))()
    lib.wc_debug_marks(buf)
    rows = [[buf[bb * 32 + k] for k in range(2)] for bb in range(148) if buf[bb * 32]]
    t0 = min(r[0] for r in rows)
    ends = [(r[1] - t0) / 1e3 for r in rows]
    starts = [(r[0] - t0) / 1e3 for r in rows]
    prof = [[buf[bb * 32 + k] for k in range(2, 8)] for bb in range(0, 148, 2) if buf[bb * 32]]
    if prof and prof[0][4]:
        n = len(prof)
        avg = [sum(r[k] for r in prof) / n for k in range(6)]
        print(f'   issuer cycles (avg over {n} leaders): wait local {avg[0]:.0f} | peer A {avg[1]:.0f} | peer B {avg[2]:.0f} | accum {avg[3]:.0f} | total {avg[4]:.0f} | k-blocks {avg[5]:.0f}')
    print(f'{tag}: start spread {max(starts):.2f} us | end min {min(ends):.1f} avg {sum(ends)/len(ends):.1f} max {max(ends):.1f}')
import os
BASE = int(os.environ.get('WC_DBG', '0'))
for flags in (1 | BASE,):
    lib.wc_debug_set(flags)
    for nm, fn in (('K3', k3), ('K6', k6)):
        for i in range(3): fn(i)
        torch.cuda.synchronize()
        fn(3); torch.cuda.synchronize()
        marks(f'flags {flags:4d} {nm}')
lib.wc_debug_set(0)
