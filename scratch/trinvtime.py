import sys, ctypes, torch
sys.path.insert(0, '/root/repo')
import wc_gan_b200 as W
from wc_gan_b200._lib import lib
x = torch.randn(64, 16, 16, 256, device='cuda') + 1
layer = W.DecorelationNormalization().cuda(); layer.train()
lib.wc_debug_set.argtypes = [ctypes.c_int]; lib.wc_debug_marks.argtypes = [ctypes.c_void_p]
import sys
lib.wc_debug_set(int(sys.argv[1]))
for _ in range(3): layer(x)
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * (160 * 32))()
lib.wc_debug_marks(buf)
for b in (0, 8):
    r = [buf[b * 32 + k] for k in range(8)]
    print('block', b, [r[k] - r[0] if r[k] else None for k in range(8)])
t0 = min(buf[b * 32] for b in range(32) if buf[b * 32])
print('start spread', [buf[b * 32] - t0 for b in range(0, 32, 4)])
