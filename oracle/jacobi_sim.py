"""TEST INFRASTRUCTURE -- an fp32 numpy model of the one-sided Jacobi iteration in wc_gan_b200/csrc/wc_jacobi.cu (round-robin
tournament, rotation threshold JROT2, stop rule JSTOP2).  The reference computes the ZCA factor with tf.svd on the host
(gan/conditional_layers.py:668-676); this file restates OUR algorithm so that its stopping rule can be checked on the CPU
against a full-convergence run and against an fp64 eigendecomposition (tests/test_jacobi_rule.py).  Nothing under
wc_gan_b200/ imports it."""
import math

import numpy as np

JROT2 = 1e-12      # rotate a pair when cos^2 > JROT2          (wc_jacobi.cu)
JSTOP2 = 1e-8      # another sweep when some pair had cos^2 > JSTOP2
JMAX_SWEEPS = 20


def rr_pairs(n, r):
    """Round-robin tournament of n (even) players, round r: rr_pair() of wc_jacobi.cu for all positions at once."""
    m = n - 1
    i = np.arange(1, n // 2)
    return np.concatenate(([m], (r + i) % m)), np.concatenate(([r], (r - i) % m))


def jacobi(L, stop2=JSTOP2, rot2=JROT2):
    """Columns of B = L V converge to U sigma.  stop2 = None: the former rule (stop after a sweep without rotations).
    Returns (B, sweeps, [max |cos| met in each sweep])."""
    B = np.array(L, dtype=np.float32)
    n = B.shape[1]
    assert n % 2 == 0
    hist = []
    for sweep in range(JMAX_SWEEPS):
        rotated, more, mx = False, False, 0.0
        for r in range(n - 1):
            P, Q = rr_pairs(n, r)
            x, y = B[:, P], B[:, Q]
            al = (x * x).sum(0, dtype=np.float32)
            be = (y * y).sum(0, dtype=np.float32)
            ga = (x * y).sum(0, dtype=np.float32)
            c2 = ga.astype(np.float64) ** 2 / (al.astype(np.float64) * be.astype(np.float64))
            mx = max(mx, float(c2.max()))
            do = ga * ga > np.float32(rot2) * al * be
            rotated |= bool(do.any())
            if stop2 is not None:
                more |= bool((ga * ga > np.float32(stop2) * al * be).any())
            gs = np.where(do, ga, np.float32(1))
            zeta = (be - al) / (2 * gs)
            tt = np.where(zeta == 0, np.float32(1), np.sign(zeta) / (np.abs(zeta) + np.sqrt(zeta * zeta + 1)))
            c = 1 / np.sqrt(tt * tt + 1)
            s = c * tt
            c = np.where(do, c, 1).astype(np.float32)
            s = np.where(do, s, 0).astype(np.float32)
            B[:, P], B[:, Q] = c * x - s * y, s * x + c * y
        hist.append(math.sqrt(mx))
        if not (more if stop2 is not None else rotated):
            break
    return B, len(hist), hist


def inverse_sqrt_from(B):
    """Sigma^-1/2 = U diag(1/sigma) U^T from the converged columns (what the library forms from U and lam)."""
    B = B.astype(np.float64)
    nn = (B * B).sum(0)
    U = B / np.sqrt(nn)
    return (U * nn ** -0.5) @ U.T
