"""debug: eager back-to-back stage times (bench.stage_times) with / without the tile-height choice of the apply kernels"""
import sys, ctypes, torch
sys.path.insert(0, '/root/repo')
import bench
import wc_gan_b200 as W
from wc_gan_b200._lib import lib
lib.wc_debug_set.argtypes = [ctypes.c_int]
dev = torch.device('cuda')
xs, dys = bench.make_inputs(torch, bench.SHAPE, 4, dev, seed=1)
for flags in (0, 4194304, 0, 4194304):
    lib.wc_debug_set(flags)
    r = bench.stage_times(torch, lib, xs, dys, 50)
    print(flags, {k: round(v * 1e3, 1) for k, v in r.items()})
