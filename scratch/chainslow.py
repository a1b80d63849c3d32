"""which CTAs of the backward chain kernel finish stage 0 last (cfg2 shape): blockIdx, stage-0 time"""
import sys, ctypes, torch, os
sys.path.insert(0, '/root/repo')
import bench, wc_gan_b200 as W
from wc_gan_b200._lib import lib
lib.wc_debug_set.argtypes = [ctypes.c_int]; lib.wc_debug_marks.argtypes = [ctypes.c_void_p]
dev = torch.device('cuda', 0)
shape, an, K, EE = (64, 16, 16, 128), os.environ.get('WC_AN', 'ufconv'), 10, 4
xs, dys = bench.make_inputs(torch, shape, 2, dev, 1)
stack = bench.build_stack(W, torch, 'd', an, shape[-1], dev, K=K, E=EE)
params = list(stack.parameters())
cls = torch.randint(0, K, (shape[0], 1), dtype=torch.int32, device=dev)
def step(i):
    x = xs[i % 2].requires_grad_(True)
    return torch.autograd.grad(stack(x, cls), [x] + params, dys[i % 2])
for i in range(3): step(i)
torch.cuda.synchronize()
lib.wc_debug_set(1)
for i in range(2): step(i)
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * (160 * 32))()
lib.wc_debug_marks(buf)
lib.wc_debug_set(0)
rows = [(b, [buf[b * 32 + k] for k in range(8, 16)]) for b in range(148) if buf[b * 32 + 8]]
t0 = min(r[0] for _, r in rows)
for k, nm in ((1, 'S0 jobs'), (3, 'S1 jobs'), (5, 'S2 jobs'), (7, 'end')):
    v = sorted(((r[k] - r[k - 1]) / 1e3, b) for b, r in rows if r[k])
    print(nm, 'slowest (us since previous mark, blockIdx):', [(round(a, 2), b) for a, b in v[-12:]])
    print('   by block:', ' '.join(f'{(r[k] - r[k-1]) / 1e3:.1f}' for b, r in rows))

# clock marks of the LAST tile job each CTA ran: 28 barrier passed, 29 entry found, 30 job descriptor in smem, 16 tile entry,
# 31 set-up done, 17 loads issued, 18.. chunk landed, 26 drained, 27 stored
names = {28: 'barrier', 29: 'found', 30: 'jobcopy', 16: 'entry', 31: 'setup', 17: 'issued', 18: 'c0', 19: 'c1', 26: 'drained', 27: 'stored'}
full = [[buf[b * 32 + k] for k in range(32)] for b in range(17)]
for b in (1, 5, 9):
    r = full[b]
    print('block', b, ' '.join(f'{names[k]}={int(r[k]) - int(r[28])}' for k in (28, 29, 30, 16, 31, 17, 18, 19, 26, 27)))
