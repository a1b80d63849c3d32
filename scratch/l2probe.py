"""How much of a streamed read set does B200's L2 keep?  Time repeated full reads of one buffer of S MB (torch sum:
read-only stream) and of a copy (read S + write S); a buffer that stays resident is re-read at L2 speed."""
import torch
dev = torch.device('cuda', 0)
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
big = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
for mb in (8, 16, 32, 48, 64, 67, 80, 96, 112, 128, 192, 256):
    x = torch.randn(mb * (1 << 20) // 4, device=dev)
    y = torch.empty_like(x)
    us_sum = t(lambda: x.sum())
    us_copy = t(lambda: y.copy_(x))
    # same read after the cache was flushed by a 512 MB write
    def cold():
        big.zero_(); return x.sum()
    us_cold = t(cold, 5) - t(lambda: big.zero_(), 5)
    print(f'{mb:4d} MB: re-read {us_sum:8.1f} us = {mb * 1.048576 / us_sum * 1e3:8.0f} GB/s | cold read {us_cold:8.1f} us = {mb * 1.048576 / max(us_cold, 1e-3) * 1e3:8.0f} GB/s | copy {us_copy:8.1f} us = {2 * mb * 1.048576 / us_copy * 1e3:8.0f} GB/s')
