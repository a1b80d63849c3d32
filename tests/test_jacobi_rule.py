"""The stop rule of the ZCA Jacobi kernels (wc_gan_b200/csrc/wc_jacobi.cu: JSTOP2), on an fp32 model of the same iteration
(oracle/jacobi_sim.py): stopping after the first sweep that meets no pair further than 1e-4 from orthogonal gives the same
inverse square root as sweeping until nothing rotates, in fewer sweeps, and both match an fp64 eigendecomposition."""
import math

import numpy as np
import pytest

from oracle import jacobi_sim as J


def covariance(kind, C, M, seed):
    rng = np.random.default_rng(seed)
    if kind == 'logspec':                       # eigenvalues log-uniform in [1e-3, 1e2], random basis
        Q, _ = np.linalg.qr(rng.standard_normal((C, C)))
        return (Q * 10 ** rng.uniform(-3, 2, C)) @ Q.T
    z = rng.standard_normal((M, C))
    if kind != 'iid':                           # oracle.make_activation: N(1, 0.5 I + 0.5 11^T), optionally through a ReLU
        z = math.sqrt(.5) * z + math.sqrt(.5) * rng.standard_normal((M, 1)) + 1
        if kind == 'relu':
            z = np.maximum(z, 0)
    z = z - z.mean(0)
    S = z.T @ z / (M - 1)
    eps = 1e-3
    return (1 - eps) * S + 2 * eps * np.eye(C)  # shrinkage + the second eps of the ZCA branch (conditional_layers.py:671,691)


@pytest.mark.parametrize('kind,C,M', [('iid', 64, 16384), ('correlated', 64, 16384), ('relu', 64, 16384), ('logspec', 64, 0),
                                      ('lowrank', 64, 48), ('correlated', 128, 8192)])
def test_stop_rule_matches_full_convergence(kind, C, M):
    S = covariance('iid' if kind == 'lowrank' else kind, C, M, seed=3)
    L = np.linalg.cholesky(S)
    B_full, n_full, hist = J.jacobi(L, stop2=None)
    B_stop, n_stop, _ = J.jacobi(L)
    lam, V = np.linalg.eigh(S)
    W_ref = (V * lam ** -0.5) @ V.T
    W_full, W_stop = J.inverse_sqrt_from(B_full), J.inverse_sqrt_from(B_stop)
    scale = np.abs(W_ref).max()
    assert n_stop <= n_full, (n_stop, n_full, hist)
    if kind in ('iid', 'correlated', 'relu'):                            # the covariances of the parity tests: at least the
        assert n_stop <= n_full - 1, (n_stop, n_full, hist)              # sweep that only verifies is saved
    assert np.abs(W_stop - W_full).max() / scale < 1e-5
    assert np.abs(W_stop - W_ref).max() / scale < 1e-4                   # the output tolerance of the layer
    # the columns the early stop leaves are orthogonal to ~1e-6: the last sweep it ran met at most 1e-4
    Bn = B_stop.astype(np.float64) / np.linalg.norm(B_stop.astype(np.float64), axis=0)
    off = np.abs(Bn.T @ Bn - np.eye(C)).max()
    assert off < 1e-5, off
