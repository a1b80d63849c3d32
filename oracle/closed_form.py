"""Closed-form (fused) statement of the WC layer that the CUDA kernels implement.

TEST INFRASTRUCTURE ONLY (see oracle/wc_oracle.py's header for the import rule).

``wc_oracle.py`` follows the reference op for op and differentiates with autograd.  This
file restates the same function in the *kernel decomposition* (SURVEY.md section 8a,
"Equivalent closed forms"): stats -> factor -> compose T_k -> apply, and the hand-derived
backward reduce -> small-matrix adjoint -> apply.  tests/test_closed_form.py checks every
output and gradient of this file against wc_oracle + autograd in fp64, so the CUDA
kernels can be checked stage by stage against the intermediates returned here.

Reference lines restated: gan/conditional_layers.py:646-712 (whitening), :790-848
(ConditionalConv11), :971-1034 (FactorizedConv11), generator.py:41-78 (composition).
"""
from __future__ import annotations

from typing import Optional

import torch

Tensor = torch.Tensor


def _blockdiag(m: Tensor) -> Tensor:
    """(G,g,g) -> (C,C)."""
    return torch.block_diag(*[m[i] for i in range(m.shape[0])])


def _diag_blocks(m: Tensor, G: int, g: int) -> Tensor:
    """(C,C) -> (G,g,g) diagonal blocks."""
    return torch.stack([m[i * g:(i + 1) * g, i * g:(i + 1) * g] for i in range(G)])


def class_weights(params: dict, after_norm: str, K: int, c: int, triangular: bool, dtype):
    """W_k (K,Cin,Cout), b_k (K,Cout), e_hat (K,E) or None: per-class coloring matrices.

    W_k = Wu + triu?(Wc_k) + sum_e e_hat[k,e] D_e   (generator.py:41-78; an Add of layer
    outputs equals the layer applied with summed matrices because all branches read the
    same whitened input)."""
    W = torch.zeros(K, c, c, dtype=dtype)
    b = torch.zeros(K, c, dtype=dtype)
    e_hat = None
    if after_norm == 'n':
        W = W + torch.eye(c, dtype=dtype)
    if after_norm in ('uconv', 'ucconv', 'ufconv'):
        W = W + params['u_kernel'].reshape(c, c)
        if params.get('u_bias') is not None:
            b = b + params['u_bias']
    if after_norm in ('cconv', 'ucconv'):
        ck = params['c_kernel'].reshape(K, c, c)
        W = W + (torch.triu(ck) if triangular else ck)
        if params.get('c_bias') is not None:
            b = b + params['c_bias']
    if after_norm in ('fconv', 'ufconv'):
        cm = params['class_matrix']
        sq = (cm * cm).sum(dim=1, keepdim=True)
        e_hat = cm * torch.rsqrt(torch.clamp(sq, min=1e-12))
        D = params['f_kernel'].reshape(-1, c, c)
        W = W + torch.einsum('ke,eij->kij', e_hat, D)
        if params.get('f_bias') is not None:
            b = b + params['f_bias']
    return W, b, e_hat


def factor(cov: Tensor, eps: float, decomposition: str):
    """cov (G,g,g) un-shrunk -> dict(Wm, sqrt, U, lam) per group (conditional_layers.py:662-678,691)."""
    G, g, _ = cov.shape
    eye = torch.eye(g, dtype=cov.dtype)
    shr = (1 - eps) * cov + eps * eye
    if decomposition == 'cholesky':
        L = torch.linalg.cholesky(shr)
        Wm = torch.linalg.solve_triangular(L, eye.expand(G, g, g), upper=False)
        return dict(Wm=Wm, sqrt=L, U=None, lam=None)
    lam, U = torch.linalg.eigh(shr + eps * eye)
    Wm = (U * lam.pow(-0.5)[:, None, :]) @ U.transpose(1, 2)
    sq = (U * lam.pow(0.5)[:, None, :]) @ U.transpose(1, 2)
    return dict(Wm=Wm, sqrt=sq, U=U, lam=lam)


def forward(x: Tensor, cls: Optional[Tensor], params: dict, *, norm='d', after_norm='uconv',
            training=True, momentum=0.99, eps=1e-3, decomposition='cholesky',
            group_size=None, triangular=False, number_of_classes=1,
            stats_override=None):
    """Returns a dict with every intermediate the kernels produce.

    stats_override = (M_total, mean (C,), S (G,g,g)) replaces the local batch
    statistics (used to emulate synchronised-batch whitening across shards)."""
    n, h, w, c = x.shape
    dt = x.dtype
    g = c if group_size is None else group_size
    G = c // g
    K = max(int(number_of_classes), 1) if after_norm in ('cconv', 'ucconv', 'fconv', 'ufconv') else 1
    M = n * h * w
    X = x.reshape(M, c)
    out = dict(M=M, G=G, g=g, K=K)
    whiten = norm in ('d', 'dr')
    renorm = norm == 'dr'
    mm = params.get('moving_mean')
    mc = params.get('moving_variance')
    if whiten:
        mc = mc.reshape(G, g, g)
        if training:
            if stats_override is None:
                mean = X.mean(dim=0)
                Xc = X - mean
                S = torch.stack([Xc[:, i * g:(i + 1) * g].t() @ Xc[:, i * g:(i + 1) * g] for i in range(G)])
                Mtot = M
            else:
                Mtot, mean, S = stats_override
            cov = S / (Mtot - 1.0)
            out['new_moving_mean'] = mm * momentum + mean.reshape(c, 1) * (1 - momentum)
            out['new_moving_variance'] = (mc * momentum + cov * (1 - momentum)).reshape(params['moving_variance'].shape)
            fb = factor(cov, eps, decomposition)
            Wm = fb['Wm']
            if renorm:
                fm = factor(mc, eps, decomposition)
                R = fm['Wm'] @ fb['sqrt']
                Weff = R @ Wm
            else:
                R, Weff = None, Wm
            out.update(M_total=Mtot, S=S, cov=cov, Wm=Wm, R=R, U=fb['U'], lam=fb['lam'])
        else:
            mean = mm.reshape(c)
            fm = factor(mc, eps, decomposition)
            Weff = fm['Wm']
            out.update(M_total=M, Wm=Weff, R=None)
    else:
        mean = torch.zeros(c, dtype=dt)
        Weff = torch.eye(g, dtype=dt).repeat(G, 1, 1)
        out.update(M_total=M)
    Wk, bk, e_hat = class_weights(params, after_norm, K, c, triangular, dt)
    Wfull = _blockdiag(Weff)
    T = torch.einsum('ji,kjl->kil', Wfull, Wk)          # T_k = Weff^T W_k
    slot = cls.reshape(n).long() if (cls is not None and K > 1) else torch.zeros(n, dtype=torch.long)
    Xc = (X - mean).reshape(n, h * w, c)
    y = Xc @ T[slot] + bk[slot][:, None, :]
    out.update(mean=mean, Weff=Weff, Wk=Wk, bk=bk, e_hat=e_hat, T=T, slot=slot,
               y=y.reshape(n, h, w, c))
    return out


def backward(x: Tensor, dy: Tensor, fwd: dict, params: dict, *, norm='d', after_norm='uconv',
             training=True, eps=1e-3, decomposition='cholesky', triangular=False,
             reduced_override=None):
    """Hand-derived backward (SURVEY.md section 8a closed forms).  Returns dict of grads and
    the intermediates (P, s, dWm, gbar, A, cvec).

    reduced_override = (dWeff (G,g,g), gbar (C,)) replaces the local reductions (emulates
    the backward all-reduce of synchronised-batch whitening)."""
    n, h, w, c = x.shape
    dt = x.dtype
    G, g, K = fwd['G'], fwd['g'], fwd['K']
    Mtot = fwd['M_total']
    whiten = norm in ('d', 'dr')
    slot = fwd['slot']
    Xc = (x.reshape(-1, c) - fwd['mean']).reshape(n, h * w, c)
    dY = dy.reshape(n, h * w, c)
    Pn = Xc.transpose(1, 2) @ dY                        # (n,C,C)
    sn = dY.sum(dim=1)                                  # (n,C)
    P = torch.zeros(K, c, c, dtype=dt).index_add_(0, slot, Pn)
    s = torch.zeros(K, c, dtype=dt).index_add_(0, slot, sn)
    Weff = fwd['Weff']
    Wfull = _blockdiag(Weff)
    Q = torch.einsum('ij,kjl->kil', Wfull, P)           # x_hat^T dY per class
    grads = {}
    if after_norm in ('uconv', 'ucconv', 'ufconv'):
        grads['u_kernel'] = Q.sum(0).reshape(1, 1, c, c)
        grads['u_bias'] = s.sum(0)
    if after_norm in ('cconv', 'ucconv'):
        grads['c_kernel'] = (torch.triu(Q) if triangular else Q).reshape(K, 1, 1, c, c)
        grads['c_bias'] = s.clone()
    if after_norm in ('fconv', 'ufconv'):
        e_hat = fwd['e_hat']
        D = params['f_kernel'].reshape(-1, c, c)
        grads['f_kernel'] = torch.einsum('ke,kij->eij', e_hat, Q).reshape(params['f_kernel'].shape)
        de = torch.einsum('kij,eij->ke', Q, D)
        cm = params['class_matrix']
        sq = (cm * cm).sum(dim=1, keepdim=True)
        rho = torch.rsqrt(torch.clamp(sq, min=1e-12))
        proj = de - e_hat * (de * e_hat).sum(dim=1, keepdim=True)
        grads['class_matrix'] = torch.where(sq > 1e-12, rho * proj, rho * de)
        if params.get('f_bias') is not None:
            grads['f_bias'] = s.clone()
    Wk = fwd['Wk']
    dWfull = torch.einsum('kij,klj->il', Wk, P)         # sum_k W_k P_k^T
    dWeff = _diag_blocks(dWfull, G, g)
    gbar = torch.einsum('kij,kj->i', Wk, s)
    if reduced_override is not None:
        dWeff, gbar = reduced_override
    T = fwd['T']
    dXc = dY @ T[slot].transpose(1, 2)
    inter = dict(P=P, s=s, Q=Q, dWeff=dWeff, gbar=gbar)
    if whiten and training:
        Wm = fwd['Wm']
        dWm = fwd['R'].transpose(1, 2) @ dWeff if fwd['R'] is not None else dWeff
        if decomposition == 'cholesky':
            Z = dWm @ Wm.transpose(1, 2)
            Phi = -(torch.tril(Z) - 0.5 * torch.diag_embed(torch.diagonal(Z, dim1=1, dim2=2)))
            Qs = Wm.transpose(1, 2) @ Phi @ Wm
            dSig = 0.5 * (Qs + Qs.transpose(1, 2))
        else:
            U, lam = fwd['U'], fwd['lam']
            r = lam.sqrt()
            F = -1.0 / (r[:, :, None] * r[:, None, :] * (r[:, :, None] + r[:, None, :]))
            sym = 0.5 * (dWm + dWm.transpose(1, 2))
            dSig = U @ ((U.transpose(1, 2) @ sym @ U) * F) @ U.transpose(1, 2)
        A = 2.0 * (1 - eps) * dSig / (Mtot - 1.0)
        cvec = (Wfull.t() @ gbar) / Mtot
        dX = dXc + Xc @ _blockdiag(A) - cvec
        inter.update(dWm=dWm, A=A, cvec=cvec)
    else:
        dX = dXc
    grads['x'] = dX.reshape(n, h, w, c)
    grads['_inter'] = inter
    return grads


# ---------------------------------------------------------------------------------------------
# Restatements of the exact algebra the fused small-matrix kernels use (checked against the
# forms above / torch.linalg in tests/test_closed_form.py)
# ---------------------------------------------------------------------------------------------
def cholesky_adjoint_psi(dWm: Tensor, Wm: Tensor, eps: float, Mtot: float) -> Tensor:
    """A = 2(1-eps) dSigma' / (M-1) as the three launches M2 / M3 / M4 of wc_small.cu compute it
    (wc_factor_bwd_impl): Z = dWm Wm^T; Psi = Phi + Phi^T built from the LOWER tiles of Z only
    (Psi_ij = -Z_max(i,j),min(i,j)); Y = Psi Wm; A = (1-eps)/(M-1) Wm^T Y.  No symmetrisation pass."""
    Z = dWm @ Wm.transpose(-1, -2)
    L = torch.tril(Z, -1)
    Psi = -(L + L.transpose(-1, -2) + torch.diag_embed(torch.diagonal(Z, dim1=-2, dim2=-1)))
    Y = Psi @ Wm
    return (1.0 - eps) / (Mtot - 1.0) * (Wm.transpose(-1, -2) @ Y)


def cluster_cholesky_inverse(A: Tensor, tb: int = 32):
    """L = chol(A) and X = L^-1 by the schedule of chol_cluster_kernel (wc_chol.cu), in plain loops:
    block row r of L left-looking, every panel as a PRODUCT with the inverse of the diagonal tile
    (L_rJ = T inv(L_JJ)^T, no triangular solve), the inverse of a diagonal tile by right-looking
    substitution on the identity, then block column r of X as X_Ir = -inv(L_II) sum_K L_IK X_Kr."""
    n = A.shape[-1]
    assert n % tb == 0
    nb = n // tb
    blk = lambda M, i, j: M[i * tb:(i + 1) * tb, j * tb:(j + 1) * tb]
    L = torch.zeros_like(A)
    X = torch.zeros_like(A)
    Dinv = []
    for r in range(nb):
        for J in range(r):
            T = blk(A, r, J).clone()
            for K in range(J):
                T -= blk(L, r, K) @ blk(L, J, K).transpose(0, 1)
            blk(L, r, J).copy_(T @ Dinv[J].transpose(0, 1))
        Td = blk(A, r, r).clone()
        for J in range(r):
            Td -= blk(L, r, J) @ blk(L, r, J).transpose(0, 1)
        # potrf of the tile, column by column; the inverse follows one column behind
        Lt = torch.zeros_like(Td)
        W = Td.clone()
        xinv = torch.eye(tb, dtype=A.dtype)            # row i = row i of L^-T (column i of L^-1)
        for k in range(tb):
            rs = 1.0 / torch.sqrt(W[k, k])
            Lt[k:, k] = W[k:, k] * rs
            W[k + 1:, k + 1:] -= torch.outer(Lt[k + 1:, k], Lt[k + 1:, k])
            xinv[:, k] *= rs
            xinv[:, k + 1:] -= torch.outer(xinv[:, k], Lt[k + 1:, k])
        blk(L, r, r).copy_(Lt)
        Dinv.append(xinv.transpose(0, 1).contiguous())   # inv(L_rr)
        blk(X, r, r).copy_(Dinv[r])
    for r in range(nb):
        for I in range(r + 1, nb):
            S = torch.zeros(tb, tb, dtype=A.dtype)
            for K in range(r, I):
                S += blk(L, I, K) @ blk(X, K, r)
            blk(X, I, r).copy_(-Dinv[I] @ S)
    return L, X


def factorised_coloring(Weff: Tensor, Wu: Optional[Tensor], D: Tensor, ehat: Tensor, P: Tensor, s: Tensor):
    """The factorised forms the chain kernel uses for fconv / ufconv with many classes (wc_small.cu, wc_bwd_chain /
    wc_compose_fwd_impl): with W_k = sum_j alpha[k,j] B_j over the E + 1 basis matrices B_0 = Wu, B_e = D_e
    (alpha[k,0] = 1, alpha[k,e] = ehat[k,e-1]; generator.py:45-48,74-76, conditional_layers.py:993-1005), no per-class
    product is ever formed:
        T_k   = sum_j alpha[k,j] (Weff^T B_j)                   G_j = sum_k alpha[k,j] P_k,  gs_j = sum_k alpha[k,j] s_k
        dWeff = sum_j B_j G_j^T     gbar = sum_j B_j gs_j       dB_j = Weff G_j  (dWu, dD)   de[k,e] = <P_k, Weff^T D_e>
    Weff (C,C) dense (block-diagonal matrices are a special case), Wu (C,C) or None, D (E,C,C), ehat (K,E), P (K,C,C) with
    P_k = sum_{n in k} xc_n^T dy_n, s (K,C).  Returns dict(T, dWeff, gbar, dWu, dD, de)."""
    K, E = ehat.shape
    one = torch.ones(K, 1, dtype=ehat.dtype) if Wu is not None else torch.zeros(K, 1, dtype=ehat.dtype)
    alpha = torch.cat([one, ehat], dim=1)                                   # (K, E+1)
    B = torch.cat([(Wu if Wu is not None else torch.zeros_like(D[0]))[None], D], dim=0)     # (E+1, C, C)
    TB = Weff.transpose(0, 1) @ B                                           # E + 1 products
    T = torch.einsum('kj,jab->kab', alpha, TB)
    G = torch.einsum('kj,kab->jab', alpha, P)
    gs = alpha.transpose(0, 1) @ s
    dWeff = torch.einsum('jab,jcb->ac', B, G)
    gbar = torch.einsum('jab,jb->a', B, gs)
    dB = Weff @ G
    de = torch.einsum('kab,eab->ke', P, TB[1:])
    return dict(T=T, dWeff=dWeff, gbar=gbar, dWu=dB[0] if Wu is not None else None, dD=dB[1:], de=de)
