import sys, torch
sys.path.insert(0, '/root/repo')
import wc_gan_b200 as W
from wc_gan_b200._lib import lib, ptr
N, HW, C = 256, 256, 256
M = N * HW
dev = torch.device('cuda')
x = torch.randn(M, C, device=dev) + 1; dy = torch.randn(M, C, device=dev)
mean = torch.zeros(C, device=dev); S = torch.zeros(C, C, device=dev)
nb = lib.wc_stats_workspace_bytes(M, C, C); ws = torch.zeros(nb // 4 + 4, device=dev)
P = torch.zeros(C * C + C, device=dev); s = P[C * C:]
st = torch.cuda.current_stream().cuda_stream
for i in range(2):
    lib.wc_stats_fwd(ptr(x), M, C, C, ptr(ws), nb, ptr(mean), ptr(S), st)
    lib.wc_reduce_bwd(ptr(x), ptr(dy), N, HW, C, C, ptr(mean), None, 1, ptr(P), ptr(s), None, 0, st)
torch.cuda.synchronize()
