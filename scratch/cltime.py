"""debug: globaltimer marks of the cluster Cholesky (one row per CTA of the cluster, ns relative to the first mark)"""
import sys, ctypes, torch
sys.path.insert(0, '/root/repo')
import wc_gan_b200 as W
from wc_gan_b200._lib import lib
lib.wc_debug_set.argtypes = [ctypes.c_int]
x = torch.randn(64, 16, 16, 256, device='cuda') + 1
layer = W.DecorelationNormalization().cuda(); layer.train()
lib.wc_debug_set(1)
for _ in range(3): layer(x)
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * (8 * 32))()
lib.wc_debug_chol_cluster_timing.argtypes = [ctypes.c_void_p]
print('rc', lib.wc_debug_chol_cluster_timing(buf))
ts = [[buf[r * 32 + i] for i in range(17)] for r in range(8)]
base = min(t[0] for t in ts if t[0])
names = ['start', 'loaded', 'waitB', 'gotB', 'trsm', 'rowdone', 'potrf0', 'potrf1', 'pubL', 'pubD', 'diagdone', 'waitA7', 'S7', 'gotD7', 'end', 'exit', 'invdone']
print('      ' + ' '.join(f'{n:>8s}' for n in names))
for r in range(8):
    print(f'cta {r} ' + ' '.join(f'{(t - base) if t else -1:8d}' for t in ts[r]))
lib.wc_debug_set(0)
