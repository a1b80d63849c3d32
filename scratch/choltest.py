"""debug: cluster Cholesky + inverse vs torch fp64, and its duration"""
import sys, ctypes, torch
sys.path.insert(0, '/root/repo')
import wc_gan_b200 as W  # noqa: F401
from wc_gan_b200._lib import lib, ptr
dev = torch.device('cuda')
st = torch.cuda.current_stream().cuda_stream
lib.wc_debug_set.argtypes = [ctypes.c_int]
torch.manual_seed(0)
for (C, g) in ((256, 256), (256, 128), (512, 256), (192, 192), (160, 160), (224, 224), (256, 64)):
    G = C // g
    Z = torch.randn(G, g, 4 * g, dtype=torch.float64)
    mix = torch.randn(G, g, g, dtype=torch.float64) / g ** 0.5 + torch.eye(g, dtype=torch.float64)
    cov = (mix @ Z) @ (mix @ Z).transpose(1, 2) / (4 * g)
    eps = 1e-3
    mm = torch.zeros(C, device=dev); mc = cov.float().to(dev).contiguous()
    Weff = torch.full((G, g, g), float('nan'), device=dev)
    info = torch.zeros(4, dtype=torch.int32, device=dev)
    nb = lib.wc_factor_workspace_bytes(C, g, 0, 0)
    ws = torch.empty(nb // 4 + 4, device=dev)
    def run():
        rc = lib.wc_factor_fwd(None, None, None, 0.0, C, g, eps, 0.99, 0, 0, 0, ptr(mm), ptr(mc), ptr(Weff), ptr(Weff), None, None, None,
                               ptr(info), ptr(ws), nb, st)
        assert rc == 0, rc
    for flags in (0, 2048):
        lib.wc_debug_set(flags)
        Weff.fill_(float('nan'))
        run(); torch.cuda.synchronize()
        A = (1 - eps) * mc.double().cpu() + eps * torch.eye(g, dtype=torch.float64)
        L = torch.linalg.cholesky(A)
        ref = torch.linalg.inv(L)
        err = float((Weff.double().cpu() - ref).abs().max() / ref.abs().max())
        for _ in range(3): run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): run()
        e1.record(); torch.cuda.synchronize()
        print(f'C={C} g={g} {"old" if flags else "new"}: rel err {err:.2e}  info {info.tolist()}  prep+factor {e0.elapsed_time(e1) / 20 * 1e3:.1f} us', flush=True)
lib.wc_debug_set(0)
# timeline of the cluster kernel (globaltimer ns relative to the earliest mark)
C = g = 256
nb = lib.wc_factor_workspace_bytes(C, g, 0, 0); ws = torch.empty(nb // 4 + 4, device=dev); info = torch.zeros(4, dtype=torch.int32, device=dev)
mm = torch.zeros(C, device=dev); Weff = torch.empty(1, g, g, device=dev)
lib.wc_debug_chol_cluster_timing.argtypes = [ctypes.c_void_p]
names = ['start', 'init', 'preB', 'gotB', 'trsm', 'rowdone', 'diagred', 'potrf', 'pubB', 'pubD', 'diagend', 'gotA7', 'part7', 'gotD7', 'end', 'exit']
Zr = torch.randn(g, 4 * g)
for tag, cov in (('equicorrelated', torch.eye(g) * 2 + 0.5), ('random', Zr @ Zr.T / (4 * g))):
    mc = cov.to(dev).contiguous()
    def run():
        lib.wc_factor_fwd(None, None, None, 0.0, C, g, 1e-3, 0.99, 0, 0, 0, ptr(mm), ptr(mc), ptr(Weff), ptr(Weff), None, None, None, ptr(info), ptr(ws), nb, st)
    for dbgf in (0, 1):
        lib.wc_debug_set(dbgf)
        for _ in range(3): run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): run()
        e1.record(); torch.cuda.synchronize()
        print(f'{tag} dbg={dbgf}: prep+factor {e0.elapsed_time(e1) / 20 * 1e3:.1f} us')
    buf = (ctypes.c_ulonglong * (8 * 32))()
    lib.wc_debug_chol_cluster_timing(buf)
    t0 = min(buf[r * 32] for r in range(8))
    print('rank ' + ' '.join(f'{n:>8s}' for n in names))
    for r in range(8):
        print(f'{r:4d} ' + ' '.join(f'{(buf[r * 32 + i] - t0) / 1e3 if buf[r * 32 + i] else 0:8.2f}' for i in range(16)))
lib.wc_debug_set(0)
