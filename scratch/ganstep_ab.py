"""cfg4 training iteration with cuDNN autotuning off / on (the convolutions are the callers of the WC path, not the path)"""
import sys, torch
sys.path.insert(0, '/root/repo')
import bench, wc_gan_b200 as W
dev = torch.device('cuda', 0)
for flag in (False, True):
    torch.backends.cudnn.benchmark = flag
    r = bench.gan_step(torch, W, dev)
    print('cudnn.benchmark', flag, round(r['ms_per_iteration'], 2), 'ms per iteration')
    torch.cuda.empty_cache()
