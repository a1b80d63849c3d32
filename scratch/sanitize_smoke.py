"""Small fwd+bwd steps of the production kernels for compute-sanitizer:
   compute-sanitizer --tool {memcheck,racecheck,synccheck} python scratch/sanitize_smoke.py
1. __graft_entry__.smoke(): 16x8x8x256 ufconv K=5 E=3 (tcgen05 K1/K3/K4/K6, cluster Cholesky with st.async hand-overs, chain)
2. 32x4x4x128 ufconv K=40 E=4: the factorised coloring stages (CJ_MIXK / CJ_MIXT / CJ_DOT2) of chain_kernel<true>
3. 8x8x8x256 uconv ZCA g=256: multi-CTA block Jacobi;  g=64: shared-memory Jacobi
4. 16x4x4x128 ucconv K=10: split class sums (EPI_ATOMIC) with the rider product Q spread over stages 0 / 1"""
import sys
sys.path.insert(0, '/root/repo')
import torch
import __graft_entry__ as G
G.smoke()
import wc_gan_b200 as W
dev = torch.device('cuda', 0)
def run(shape, an, K, E, dec='cholesky', g=None):
    st = W.WhiteningColoring('d', an, number_of_classes=K, filters_emb=E, name='s', decomposition=dec, group_size=g)
    st.build(shape[-1], dev); st.built = True
    st.to(dev).train()
    x = (torch.randn(shape, device=dev) + 1).requires_grad_(True)
    cls = torch.randint(0, K or 1, (shape[0], 1), dtype=torch.int32, device=dev)
    y = st(x, cls)
    y.backward(torch.randn_like(y))
    torch.cuda.synchronize()
    st.check_info()
    assert torch.isfinite(y).all() and torch.isfinite(x.grad).all()
run((32, 4, 4, 128), 'ufconv', 40, 4)
run((16, 4, 4, 128), 'ucconv', 10, 4)
run((8, 8, 8, 256), 'uconv', None, 4, dec='zca')
run((8, 8, 8, 256), 'uconv', None, 4, dec='zca', g=64)
print('sanitize_smoke ok')
