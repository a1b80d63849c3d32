"""oracle/closed_form.py (the kernel decomposition + hand-derived backward) against
oracle/wc_oracle.py (reference op-for-op + autograd), fp64.  No reference test pins
any gradient, so this is the pin for the backward formulas."""
import itertools

import pytest
import torch

from oracle import closed_form as CF
from oracle import wc_oracle as O

DT = torch.float64


def _run(norm, after_norm, dec, g, training, triangular=False, K=3, E=2, c=8, shape=(5, 3, 2)):
    n, h, w = shape
    x = O.make_activation((n, h, w, c), 'correlated', seed=1, dtype=DT)
    gen = torch.Generator().manual_seed(2)
    cls = torch.randint(0, K, (n, 1), generator=gen, dtype=torch.int32)
    params = O.make_params(c, after_norm=after_norm, number_of_classes=K, filters_emb=E,
                           group_size=g, seed=3, dtype=DT, random_bias=True, random_moving=True)
    dy = torch.randn(n, h, w, c, generator=gen, dtype=DT)
    leaves = {k: v.clone().requires_grad_(True) for k, v in params.items()
              if k not in ('moving_mean', 'moving_variance')}
    p2 = dict(params)
    p2.update(leaves)
    xr = x.clone().requires_grad_(True)
    y, mm, mc = O.wc_block(xr, cls, p2, norm=norm, after_norm=after_norm, training=training,
                           decomposition=dec, group_size=g, triangular=triangular)
    grads = torch.autograd.grad(y, [xr] + list(leaves.values()), dy, allow_unused=True)
    fwd = CF.forward(x, cls, params, norm=norm, after_norm=after_norm, training=training,
                     decomposition=dec, group_size=g, triangular=triangular, number_of_classes=K)
    bwd = CF.backward(x, dy, fwd, params, norm=norm, after_norm=after_norm, training=training,
                      decomposition=dec, triangular=triangular)
    assert torch.allclose(fwd['y'], y, rtol=1e-9, atol=1e-9)
    if training and norm != 'n':
        assert torch.allclose(fwd['new_moving_mean'], mm, rtol=1e-12, atol=1e-12)
        assert torch.allclose(fwd['new_moving_variance'], mc, rtol=1e-12, atol=1e-12)
    scale = grads[0].abs().max()
    assert (bwd['x'] - grads[0]).abs().max() <= 1e-8 * scale
    for name, gr in zip(leaves.keys(), grads[1:]):
        if gr is None:
            continue
        assert name in bwd, name
        assert (bwd[name].reshape(gr.shape) - gr).abs().max() <= 1e-8 * max(gr.abs().max(), 1e-30), name


@pytest.mark.parametrize('after_norm', ['uconv', 'cconv', 'ucconv', 'fconv', 'ufconv', 'n'])
@pytest.mark.parametrize('dec', ['cholesky', 'zca'])
@pytest.mark.parametrize('g', [None, 4])
def test_train(after_norm, dec, g):
    _run('d', after_norm, dec, g, True)


@pytest.mark.parametrize('after_norm', ['uconv', 'ucconv', 'ufconv'])
@pytest.mark.parametrize('dec', ['cholesky', 'zca'])
def test_renorm(after_norm, dec):
    _run('dr', after_norm, dec, None, True)
    _run('dr', after_norm, dec, 2, True)


@pytest.mark.parametrize('after_norm', ['uconv', 'cconv', 'ufconv'])
def test_inference(after_norm):
    _run('d', after_norm, 'cholesky', None, False)
    _run('d', after_norm, 'zca', 4, False)


@pytest.mark.parametrize('after_norm', ['uconv', 'cconv', 'ucconv', 'fconv', 'ufconv'])
def test_coloring_only(after_norm):
    _run('n', after_norm, 'cholesky', None, True)


def test_triangular():
    _run('d', 'cconv', 'cholesky', None, True, triangular=True)
    _run('d', 'ucconv', 'cholesky', 4, True, triangular=True)


def test_gradcheck_oracle():
    """autograd.gradcheck on the op-for-op oracle itself."""
    c = 4
    x = O.make_activation((3, 2, 2, c), 'correlated', seed=5, dtype=DT).requires_grad_(True)
    mm, mc = torch.zeros(c, 1, dtype=DT), torch.eye(c, dtype=DT)
    for dec in ('cholesky', 'zca'):
        f = lambda t: O.decorelation_normalization(t, mm, mc, training=True, decomposition=dec)[0]
        assert torch.autograd.gradcheck(f, (x,), eps=1e-6, atol=1e-5, rtol=1e-4)


def test_sync_batch_emulation():
    """Chan-merged shard statistics reproduce the full-batch layer."""
    c, g = 8, 4
    x = O.make_activation((6, 3, 2, c), 'relu', seed=7, dtype=DT)
    params = O.make_params(c, after_norm='uconv', group_size=g, seed=1, dtype=DT)
    full = CF.forward(x, None, params, group_size=g)
    shards = [x[:2], x[2:]]
    stats = []
    for sh in shards:
        X = sh.reshape(-1, c)
        mu = X.mean(0)
        Xc = X - mu
        S = torch.stack([Xc[:, i * g:(i + 1) * g].t() @ Xc[:, i * g:(i + 1) * g] for i in range(c // g)])
        stats.append((X.shape[0], mu, S))
    n = sum(s[0] for s in stats)
    mu = sum(s[0] * s[1] for s in stats) / n
    S = torch.zeros_like(stats[0][2])
    for (nr, mur, Sr) in stats:
        d = (mur - mu).reshape(c // g, g)
        S = S + Sr + nr * d[:, :, None] * d[:, None, :]
    ys = [CF.forward(sh, None, params, group_size=g, stats_override=(n, mu, S))['y'] for sh in shards]
    assert torch.allclose(torch.cat(ys), full['y'], rtol=1e-10, atol=1e-10)


def test_psi_form_of_the_cholesky_adjoint():
    """The M2 / M3 / M4 launches build A from Psi = Phi + Phi^T (lower tiles of Z only); same A as sym(Wm^T Phi Wm)."""
    torch.manual_seed(0)
    G, g, eps, M = 3, 8, 1e-3, 77.0
    S = torch.randn(G, g, 2 * g, dtype=torch.float64)
    cov = S @ S.transpose(1, 2) / (2 * g) + 0.1 * torch.eye(g, dtype=torch.float64)
    Wm = torch.linalg.inv(torch.linalg.cholesky(cov))
    dWm = torch.randn(G, g, g, dtype=torch.float64)
    Z = dWm @ Wm.transpose(1, 2)
    Phi = -(torch.tril(Z) - 0.5 * torch.diag_embed(torch.diagonal(Z, dim1=1, dim2=2)))
    Qs = Wm.transpose(1, 2) @ Phi @ Wm
    ref = 2.0 * (1 - eps) * 0.5 * (Qs + Qs.transpose(1, 2)) / (M - 1.0)
    got = CF.cholesky_adjoint_psi(dWm, Wm, eps, M)
    assert float((got - ref).abs().max()) < 1e-12 * float(ref.abs().max()) + 1e-15


@pytest.mark.parametrize('n', [128, 256])
def test_cluster_cholesky_schedule(n):
    """The cluster kernel's schedule (panels as products with the inverse of the diagonal tile, inverse one column
    behind the factorisation, block columns of L^-1) gives L and L^-1."""
    torch.manual_seed(1)
    S = torch.randn(n, 2 * n, dtype=torch.float64)
    A = S @ S.t() / (2 * n) + 1e-3 * torch.eye(n, dtype=torch.float64)
    L, X = CF.cluster_cholesky_inverse(A)
    Lr = torch.linalg.cholesky(A)
    assert float((L - Lr).abs().max()) < 1e-10 * float(Lr.abs().max())
    assert float((X @ Lr - torch.eye(n, dtype=torch.float64)).abs().max()) < 1e-8
    assert float(torch.triu(X, 1).abs().max()) == 0.0


@pytest.mark.parametrize('has_u', [True, False])
def test_factorised_coloring_forms(has_u):
    """fconv / ufconv with many classes: the E + 1 basis products + mixing passes of the chain kernel equal the direct
    per-class forms (T_k = Weff^T W_k, dWeff = sum_k W_k P_k^T, gbar = sum_k W_k s_k, dW_k = Weff P_k reduced to dWu / dD /
    de) that the rest of the library -- and autograd over the oracle -- use."""
    torch.manual_seed(2)
    C, K, E = 12, 17, 3
    f = lambda *sh: torch.randn(*sh, dtype=torch.float64)
    Weff, Wu, D, P, s = f(C, C), (f(C, C) if has_u else None), f(E, C, C), f(K, C, C), f(K, C)
    ehat = torch.nn.functional.normalize(f(K, E), dim=1)
    got = CF.factorised_coloring(Weff, Wu, D, ehat, P, s)
    Wk = torch.einsum('ke,eab->kab', ehat, D) + (Wu if has_u else 0.0)
    Q = Weff @ P                                                            # dW_k
    ref = dict(T=Weff.t() @ Wk, dWeff=torch.einsum('kab,kcb->ac', Wk, P), gbar=torch.einsum('kab,kb->a', Wk, s),
               dWu=Q.sum(0) if has_u else None, dD=torch.einsum('ke,kab->eab', ehat, Q), de=torch.einsum('kab,eab->ke', Q, D))
    for name, r in ref.items():
        if r is None:
            assert got[name] is None
            continue
        assert float((got[name] - r).abs().max()) < 1e-12 * float(r.abs().max()) + 1e-14, name
