"""The fine-grained C ABI (the stage functions INTEGRATION.md section 3 tells a maintainer to bind one by one:
wc_stats_fwd -> wc_factor_fwd -> wc_compose_fwd -> [wc_pack_fwd] -> wc_apply_fwd and wc_reduce_bwd -> wc_color_bwd ->
wc_factor_bwd -> [wc_pack_bwd] -> wc_apply_bwd) driven directly through ctypes with raw device pointers, against the
fp64 oracle.  Tolerances: outputs <= 1e-4, gradients <= 1e-3 of max-abs (north_star)."""
import pytest
import torch

from oracle import wc_oracle as O

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def rel_err(a, b):
    a = a.detach().double().cpu().reshape(-1)
    b = b.detach().double().cpu().reshape(-1)
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-30))


def _f(*shape):
    return torch.zeros(shape, dtype=torch.float32, device=DEV)


@pytest.mark.parametrize('shape,after_norm,K,E,g,dec', [
    ((16, 8, 8, 256), 'uconv', 1, 0, 256, 'cholesky'),      # tcgen05 kernels, cluster Cholesky
    ((12, 6, 6, 128), 'ucconv', 4, 0, 128, 'cholesky'),     # tcgen05 kernels, per-class matrices
    ((6, 5, 5, 48), 'ufconv', 3, 2, 16, 'cholesky'),        # FFMA kernels, grouped whitening, soft assignment
    ((8, 4, 4, 64), 'uconv', 1, 0, 32, 'zca'),              # ZCA factor and its adjoint
])
def test_stage_functions_against_oracle(shape, after_norm, K, E, g, dec):
    from wc_gan_b200._lib import lib, ptr, WC_DECOMP
    N, H, Wd, C = shape
    HW, M, G = H * Wd, N * H * Wd, C // g
    eps, mom = 1e-3, 0.99
    st = torch.cuda.current_stream().cuda_stream
    x32 = O.make_activation(shape, 'correlated', seed=2)
    gen = torch.Generator().manual_seed(7)
    cls = torch.randint(0, max(K, 1), (N, 1), generator=gen, dtype=torch.int32)
    dy32 = torch.randn(shape, generator=gen)
    params = O.make_params(C, after_norm=after_norm, number_of_classes=K, filters_emb=E, group_size=None if g == C else g,
                           seed=5, random_bias=True, random_moving=True)

    # ---- oracle (fp64)
    p64 = {k: v.double() for k, v in params.items()}
    leaves = {k: v.clone().requires_grad_(True) for k, v in p64.items() if k not in ('moving_mean', 'moving_variance')}
    p64.update(leaves)
    x64 = x32.double().requires_grad_(True)
    y_ref, mm_ref, mc_ref = O.wc_block(x64, cls, p64, norm='d', after_norm=after_norm, training=True, decomposition=dec,
                                       group_size=None if g == C else g)
    gr = dict(zip(['x'] + list(leaves), torch.autograd.grad(y_ref, [x64] + list(leaves.values()), dy32.double(),
                                                            allow_unused=True)))

    # ---- device operands, reference weight layouts flattened to what include/wc_b200.h documents
    x, dy = x32.to(DEV), dy32.to(DEV)
    clsd = cls.reshape(-1).to(DEV) if K > 1 else None
    Wu = params['u_kernel'].reshape(C, C).contiguous().to(DEV)
    bu = params['u_bias'].to(DEV)
    Wc = bc = cm = D = None
    if after_norm == 'ucconv':
        Wc, bc = params['c_kernel'].reshape(K, C, C).contiguous().to(DEV), params['c_bias'].contiguous().to(DEV)
    if after_norm == 'ufconv':
        cm, D = params['class_matrix'].contiguous().to(DEV), params['f_kernel'].reshape(E, C, C).contiguous().to(DEV)
    mmean = params['moving_mean'].reshape(-1).clone().to(DEV)
    mcov = params['moving_variance'].clone().contiguous().to(DEV)
    decomp = WC_DECOMP[dec]
    tc = bool(lib.wc_path_for(M, C, C, g) & 1)

    # ---- forward: K1, K2, K2b, (pack), K3
    mean, S = _f(C), _f(G, g, g)
    nb = lib.wc_stats_workspace_bytes(M, C, g)
    ws1 = _f(nb // 4 + 64)
    assert lib.wc_stats_fwd(ptr(x), M, C, g, ptr(ws1), nb, ptr(mean), ptr(S), st) == 0
    Weff, U, lam = _f(G, g, g), _f(G, g, g), _f(C)
    info = torch.zeros(4, dtype=torch.int32, device=DEV)
    nf = lib.wc_factor_workspace_bytes(C, g, decomp, 0)
    ws2 = _f(nf // 4 + 64)
    assert lib.wc_factor_fwd(ptr(mean), ptr(S), None, float(M), C, g, eps, mom, decomp, 0, 1, ptr(mmean), ptr(mcov),
                             ptr(Weff), ptr(Weff), None, ptr(U), ptr(lam), ptr(info), ptr(ws2), nf, st) == 0
    ehat, Wk, T, b = _f(max(K, 1), max(E, 1)), _f(K, C, C), _f(K, C, C), _f(K, C)
    assert lib.wc_compose_fwd(ptr(Weff), C, C, g, K, ptr(Wu), ptr(bu), ptr(Wc), ptr(bc), 0, ptr(cm), ptr(D), E, None,
                              ptr(ehat), ptr(Wk), ptr(T), ptr(b), st) == 0
    Tp = None
    if tc:
        Tp = _f(lib.wc_pack_bytes(C, K, 0) // 4)
        assert lib.wc_pack_fwd(ptr(T), C, K, ptr(Tp), st) == 0
    y = torch.empty_like(x)
    assert lib.wc_apply_fwd(ptr(x), N, HW, C, C, ptr(mean), ptr(T), ptr(Tp), ptr(b), ptr(clsd), K, 0, ptr(y), ptr(info), st) == 0
    torch.cuda.synchronize()
    assert info.tolist()[:2] == [0, 0]
    assert rel_err(y, y_ref) <= 1e-4
    assert rel_err(mmean, mm_ref) <= 1e-5 and rel_err(mcov, mc_ref) <= 1e-5

    # ---- backward: K4, K5a, K5b, (pack), K6
    P, s = _f(K, C, C), _f(K, C)
    nr = lib.wc_reduce_workspace_bytes(N, HW, C, C, K)
    ws3 = _f(nr // 4 + 64)
    assert lib.wc_reduce_bwd(ptr(x), ptr(dy), N, HW, C, C, ptr(mean), ptr(clsd), K, ptr(P), ptr(s), ptr(ws3), nr, st) == 0
    dWu, dbu = _f(C, C), _f(C)
    dWc = _f(K, C, C) if Wc is not None else None
    dbc = _f(K, C) if bc is not None else None
    dD = _f(E, C, C) if D is not None else None
    dcm = _f(K, E) if cm is not None else None
    dWeff, gbar = _f(G, g, g), _f(C)
    nc = lib.wc_color_bwd_workspace_bytes(C, C, g, K, E)
    ws4 = _f(nc // 4 + 64)
    assert lib.wc_color_bwd(ptr(P), ptr(s), ptr(Weff), ptr(Wk), C, C, g, K, ptr(dWu), ptr(dbu), ptr(dWc), ptr(dbc), 0,
                            ptr(cm), ptr(ehat), ptr(D), E, ptr(dD), ptr(dcm), None, ptr(dWeff), ptr(gbar),
                            ptr(ws4), nc, st) == 0
    A, cvec = _f(G, g, g), _f(C)
    nfb = lib.wc_factor_bwd_workspace_bytes(C, g, decomp, 0)
    ws5 = _f(nfb // 4 + 64)
    assert lib.wc_factor_bwd(ptr(dWeff), ptr(gbar), ptr(Weff), ptr(Weff), None, ptr(U), ptr(lam), None, float(M), C, g,
                             eps, decomp, 0, ptr(A), ptr(cvec), ptr(ws5), nfb, st) == 0
    TAp = None
    if tc:
        TAp = _f(lib.wc_pack_bytes(C, K, 1) // 4)
        assert lib.wc_pack_bwd(ptr(T), ptr(A), C, g, K, ptr(TAp), st) == 0
    dx = torch.empty_like(x)
    assert lib.wc_apply_bwd(ptr(x), ptr(dy), N, HW, C, C, ptr(mean), ptr(T), ptr(A), ptr(TAp), ptr(cvec), g, ptr(clsd), K,
                            ptr(dx), st) == 0
    torch.cuda.synchronize()
    assert rel_err(dx, gr['x']) <= 1e-3
    assert rel_err(dWu, gr['u_kernel']) <= 1e-3 and rel_err(dbu, gr['u_bias']) <= 1e-3
    if dWc is not None:
        assert rel_err(dWc, gr['c_kernel']) <= 1e-3 and rel_err(dbc, gr['c_bias']) <= 1e-3
    if dD is not None:
        assert rel_err(dD, gr['f_kernel']) <= 1e-3 and rel_err(dcm, gr['class_matrix']) <= 1e-3
