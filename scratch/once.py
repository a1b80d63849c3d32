"""one launch sequence K1, K3, K4, K6 (x3) on the headline shape, for ncu captures"""
import sys, torch
sys.path.insert(0, '/root/repo')
import wc_gan_b200 as W  # noqa: F401
from wc_gan_b200._lib import lib, ptr
N, HW, C = 256, 256, 256
M = N * HW
dev = torch.device('cuda')
xs = [torch.randn(M, C, device=dev) + 1 for _ in range(2)]
dys = [torch.randn(M, C, device=dev) for _ in range(2)]
out = torch.empty(M, C, device=dev)
mean = torch.zeros(C, device=dev); S = torch.zeros(C, C, device=dev)
nb = lib.wc_stats_workspace_bytes(M, C, C); ws = torch.zeros(nb // 4 + 4, device=dev)
P = torch.zeros(C * C + C, device=dev); s = P[C * C:]
T = torch.randn(C, C, device=dev) / 16; A = torch.randn(C, C, device=dev) / 1e4; b = torch.randn(C, device=dev)
st = torch.cuda.current_stream().cuda_stream
pf = torch.empty(lib.wc_pack_bytes(C, 1, 0) // 4, device=dev); pb = torch.empty(lib.wc_pack_bytes(C, 1, 1) // 4, device=dev)
lib.wc_pack_fwd(ptr(T), C, 1, ptr(pf), st); lib.wc_pack_bwd(ptr(T), ptr(A), C, C, 1, ptr(pb), st)
info = torch.zeros(4, dtype=torch.int32, device=dev)
for i in range(3):
    x, dy = xs[i % 2], dys[i % 2]
    assert lib.wc_stats_fwd(ptr(x), M, C, C, ptr(ws), nb, ptr(mean), ptr(S), st) == 0
    assert lib.wc_apply_fwd(ptr(x), N, HW, C, C, ptr(mean), ptr(T), ptr(pf), ptr(b), None, 1, 0, ptr(out), ptr(info), st) == 0
    assert lib.wc_reduce_bwd(ptr(x), ptr(dy), N, HW, C, C, ptr(mean), None, 1, ptr(P), ptr(s), None, 0, st) == 0
    assert lib.wc_apply_bwd(ptr(x), ptr(dy), N, HW, C, C, ptr(mean), ptr(T), ptr(A), ptr(pb), ptr(b), C, None, 1, ptr(out), st) == 0
torch.cuda.synchronize()
