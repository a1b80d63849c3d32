"""debug: pinned-memory copy bandwidth of this box (what bounds bench.py's e2e leg)"""
import torch, time, os
n = 64 * 1024 * 1024
d = torch.empty(n, dtype=torch.float32, device='cuda')
h = torch.empty(n, dtype=torch.float32, pin_memory=True); h.fill_(1.0)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
d2 = torch.empty_like(d); h2 = torch.empty(n, dtype=torch.float32, pin_memory=True); h2.fill_(0.0)
def t(fn, k=5):
    fn(); torch.cuda.synchronize()
    a = time.time()
    for _ in range(k): fn()
    torch.cuda.synchronize()
    return (time.time() - a) / k
gb = n * 4 / 1e9
print('cpus', len(os.sched_getaffinity(0)), 'H2D GB/s', round(gb / t(lambda: d.copy_(h, non_blocking=True)), 1),
      'D2H GB/s', round(gb / t(lambda: h2.copy_(d2, non_blocking=True)), 1))
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
print('duplex GB/s each way', round(gb / t(both), 1))
