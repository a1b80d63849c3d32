"""debug: time K1 / K4 alone and print in-kernel globaltimer marks"""
import sys, ctypes, torch
sys.path.insert(0, '/root/repo')
import wc_gan_b200 as W
from wc_gan_b200._lib import lib, ptr
N, HW, C = 256, 256, 256
M = N * HW
dev = torch.device('cuda')
xs = [torch.randn(M, C, device=dev) + 1 for _ in range(3)]
dys = [torch.randn(M, C, device=dev) for _ in range(3)]
mean = torch.zeros(C, device=dev); S = torch.zeros(C, C, device=dev)
nb = lib.wc_stats_workspace_bytes(M, C, C); ws = torch.zeros(nb // 4 + 4, device=dev)
P = torch.zeros(C * C + C, device=dev); s = P[C * C:]
st = torch.cuda.current_stream().cuda_stream
lib.wc_debug_set.argtypes = [ctypes.c_int]; lib.wc_debug_marks.argtypes = [ctypes.c_void_p]
def marks(tag):
    buf = (ctypes.c_ulonglong * (160 * 32))()
    lib.wc_debug_marks(buf)
    rows = [[buf[b * 32 + k] for k in range(3)] for b in range(148) if buf[b * 32]]
    t0 = min(r[0] for r in rows)
    main = [r[1] - r[0] for r in rows]; fl = [r[2] - r[1] for r in rows]; end = [r[2] - t0 for r in rows]
    print(f'{tag}: ctas {len(rows)} main avg {sum(main)/len(main)/1e3:.1f} max {max(main)/1e3:.1f} us | flush avg {sum(fl)/len(fl)/1e3:.1f} max {max(fl)/1e3:.1f} us | end max {max(end)/1e3:.1f} us, start spread {max(r[0] for r in rows)-t0} ns')
def timeit(fn, n=20):
    for i in range(3): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
k1 = lambda i: lib.wc_stats_fwd(ptr(xs[i % 3]), M, C, C, ptr(ws), nb, ptr(mean), ptr(S), st)
k4 = lambda i: lib.wc_reduce_bwd(ptr(xs[i % 3]), ptr(dys[i % 3]), N, HW, C, C, ptr(mean), None, 1, ptr(P), ptr(s), None, 0, st)
for flags in (1, 5, 9, 13, 3):
    lib.wc_debug_set(flags)
    print('flags', flags, 'K1 stage us', round(timeit(k1), 1), 'K4 stage us', round(timeit(k4), 1))
    if flags & 1:
        k1(0); torch.cuda.synchronize(); marks('K1')
        k4(0); torch.cuda.synchronize(); marks('K4')
