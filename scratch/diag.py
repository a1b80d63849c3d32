"""debug: which pipeline stage bounds the tcgen05 kernels?  Times K1/K3/K4/K6 alone with parts switched off
(wc_debug_set experiment bits; results are wrong by construction, only the durations matter)."""
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
mean = torch.zeros(C, device=dev); S = torch.zeros(C, C, device=dev)
nb = lib.wc_stats_workspace_bytes(M, C, C); ws = torch.zeros(nb // 4 + 4, device=dev)
P = torch.zeros(C * C + C, device=dev); s = P[C * C:]
T = torch.randn(C, C, device=dev) / 16; A = torch.randn(C, C, device=dev) / 1e4; b = torch.randn(C, device=dev)
st = torch.cuda.current_stream().cuda_stream
lib.wc_debug_set.argtypes = [ctypes.c_int]; lib.wc_debug_marks.argtypes = [ctypes.c_void_p]
pf = torch.empty(lib.wc_pack_bytes(C, 1, 0) // 4, device=dev); pb = torch.empty(lib.wc_pack_bytes(C, 1, 1) // 4, device=dev)
assert lib.wc_pack_fwd(ptr(T), C, 1, ptr(pf), st) == 0
assert lib.wc_pack_bwd(ptr(T), ptr(A), C, C, 1, ptr(pb), st) == 0
info = torch.zeros(4, dtype=torch.int32, device=dev)


def timeit(fn, n=20):
    for i in range(4): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def chk(rc):
    assert rc == 0, rc


k1 = lambda i: chk(lib.wc_stats_fwd(ptr(xs[i % NB]), M, C, C, ptr(ws), nb, ptr(mean), ptr(S), st))
k4 = lambda i: chk(lib.wc_reduce_bwd(ptr(xs[i % NB]), ptr(dys[i % NB]), N, HW, C, C, ptr(mean), None, 1, ptr(P), ptr(s), None, 0, st))
k3 = lambda i: chk(lib.wc_apply_fwd(ptr(xs[i % NB]), N, HW, C, C, ptr(mean), ptr(T), ptr(pf), ptr(b), None, 1, 0, ptr(outs[i % 2]), ptr(info), st))
k6 = lambda i: chk(lib.wc_apply_bwd(ptr(xs[i % NB]), ptr(dys[i % NB]), N, HW, C, C, ptr(mean), ptr(T), ptr(A), ptr(pb), ptr(b), C, None, 1, ptr(outs[i % 2]), st))

print('gram kernels: bit2 no red.global, bit4 no MMA, bit8 no convert')
for flags in (0, 16, 2, 4, 8, 2 | 4, 4 | 8, 2 | 4 | 8):
    lib.wc_debug_set(flags)
    print(f'  flags {flags:5d}  K1 {timeit(k1):6.1f} us   K4 {timeit(k4):6.1f} us', flush=True)
print('apply kernels: 64 no B reload, 128 no convert, 256 no MMA, 512 no store, 1024 no A load')
for flags in (0, 32768, 4096, 64, 128, 256, 512, 1024, 64 | 256, 64 | 128 | 256, 64 | 128 | 256 | 512, 64 | 128 | 256 | 1024, 64 | 128 | 256 | 512 | 1024,
              128 | 256 | 512 | 1024):
    lib.wc_debug_set(flags)
    print(f'  flags {flags:5d}  K3 {timeit(k3):6.1f} us   K6 {timeit(k6):6.1f} us', flush=True)
lib.wc_debug_set(0)
# plain copy of the same tensor for reference (read 67 MB + write 67 MB)
cp = lambda i: outs[i % 2].copy_(xs[i % NB])
print(f'torch copy_ 67 MB -> 67 MB: {timeit(cp):6.1f} us')
