import sys, ctypes, torch
sys.path.insert(0, '/root/repo')
import wc_gan_b200 as W
from wc_gan_b200._lib import lib
x = torch.randn(64, 16, 16, 256, device='cuda') + 1
layer = W.DecorelationNormalization().cuda(); layer.train()
for _ in range(3): layer(x)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 64)()
lib.wc_debug_chol_timing.argtypes = [ctypes.c_void_p]
print('rc', lib.wc_debug_chol_timing(buf))
ts = list(buf)
base = ts[0]
for J in range(8):
    a, b, c = ts[1+3*J], ts[2+3*J], ts[3+3*J]
    nxt = ts[1+3*(J+1)] if J < 7 else ts[30]
    print(f'J={J}: potrf {b-a}  trsm+inv {c-b}  update {nxt-c}')
print('load', ts[0]-ts[62], 'chol', ts[30]-ts[0], 'Lout', ts[63]-ts[30], 'total', ts[63]-ts[62])
