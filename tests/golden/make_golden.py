"""Generates tests/golden/wc_golden.npz: frozen input/output vectors of the WC layer.

The reference (TF 1.5 / Keras 2.0.8, Python 2) cannot be imported in this image, so the vectors come from
oracle/wc_oracle.py in fp64 (the op-for-op restatement pinned by the reference's own known-answer tests,
tests/test_oracle_reference_vectors.py).  They freeze the oracle: tests/test_golden.py checks that the oracle still
reproduces them (CPU) and that the CUDA path matches them within the parity tolerances (GPU).

    python tests/golden/make_golden.py          # rewrites the fixture (only when the oracle is changed on purpose)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import wc_oracle as O  # noqa: E402

# (name, norm, after_norm, shape, decomposition, group_size, K, E, training)
CASES = [
    ('chol_uconv', 'd', 'uconv', (6, 4, 4, 16), 'cholesky', None, 1, 0, True),
    ('chol_ucconv', 'd', 'ucconv', (6, 3, 5, 12), 'cholesky', None, 4, 0, True),
    ('zca_ufconv_g4', 'd', 'ufconv', (5, 4, 4, 8), 'zca', 4, 3, 2, True),
    ('renorm_cconv', 'dr', 'cconv', (6, 4, 4, 8), 'cholesky', None, 3, 0, True),
    ('infer_fconv', 'd', 'fconv', (4, 3, 3, 8), 'cholesky', None, 3, 2, False),
    ('tc_shape_chol_uconv', 'd', 'uconv', (3, 8, 8, 128), 'cholesky', None, 1, 0, True),
]


def run(case):
    name, norm, an, shape, dec, g, K, E, training = case
    x32 = O.make_activation(shape, 'correlated', seed=11)
    gen = torch.Generator().manual_seed(12)
    cls = torch.randint(0, max(K, 1), (shape[0], 1), generator=gen, dtype=torch.int32)
    dy32 = torch.randn(shape, generator=gen)
    params = O.make_params(shape[-1], after_norm=an, number_of_classes=K, filters_emb=E, group_size=g, seed=13,
                           random_bias=True, random_moving=True)
    p64 = {k: v.double() for k, v in params.items()}
    x64 = x32.double().requires_grad_(True)
    y, mm, mc = O.wc_block(x64, cls, p64, norm=norm, after_norm=an, training=training, decomposition=dec, group_size=g)
    dx, = torch.autograd.grad(y, [x64], dy32.double())
    out = {'x': x32.numpy(), 'cls': cls.numpy(), 'dy': dy32.numpy(), 'y': y.detach().numpy(), 'dx': dx.numpy(),
           'moving_mean': mm.detach().numpy(), 'moving_variance': mc.detach().numpy()}
    for k, v in params.items():
        out['p_' + k] = v.numpy()
    return {f'{name}/{k}': v for k, v in out.items()}


def main():
    blob = {}
    for c in CASES:
        blob.update(run(c))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'wc_golden.npz')
    np.savez_compressed(path, **blob)
    print(path, os.path.getsize(path), 'bytes,', len(blob), 'arrays')


if __name__ == '__main__':
    main()
