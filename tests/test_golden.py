"""Frozen golden vectors (tests/golden/wc_golden.npz, written by tests/golden/make_golden.py from the fp64 oracle).
CPU: the oracle still reproduces them.  GPU: the CUDA path matches them (outputs 1e-4, input gradient 1e-3, relative to
max-abs), through the reference-facing factory."""
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch

from oracle import wc_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location('make_golden', os.path.join(HERE, 'golden', 'make_golden.py'))
MG = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(MG)
GOLD = np.load(os.path.join(HERE, 'golden', 'wc_golden.npz'))


def _case(name):
    pre = name + '/'
    return {k[len(pre):]: torch.from_numpy(GOLD[k]) for k in GOLD.files if k.startswith(pre)}


@pytest.mark.parametrize('case', MG.CASES, ids=[c[0] for c in MG.CASES])
def test_oracle_reproduces_golden(case):
    fresh = MG.run(case)
    for k, v in fresh.items():
        ref = GOLD[k]
        assert v.shape == ref.shape, k
        if v.dtype.kind in 'iu':
            assert np.array_equal(v, ref), k
        else:
            assert np.max(np.abs(v - ref)) <= 1e-9 * max(1.0, float(np.max(np.abs(ref)))), k


@pytest.mark.gpu
@pytest.mark.parametrize('case', MG.CASES, ids=[c[0] for c in MG.CASES])
def test_cuda_matches_golden(case):
    import wc_gan_b200 as W
    sys.path.insert(0, HERE)
    from test_gpu_parity import load_params, rel_err       # the same loaders the parity tests use
    name, norm, an, shape, dec, g, K, E, training = case
    d = _case(name)
    params = {k[2:]: v for k, v in d.items() if k.startswith('p_')}
    stack = W.WhiteningColoring(norm, an, number_of_classes=K, filters_emb=E, name='g', decomposition=dec, group_size=g)
    stack.build(shape[-1], torch.device('cuda')); stack.built = True
    stack.to('cuda')
    load_params(stack, params)
    stack.train(training)
    x = d['x'].cuda().requires_grad_(True)
    y = stack(x, d['cls'].cuda())
    y.backward(d['dy'].cuda())
    torch.cuda.synchronize()
    stack.check_info()
    assert rel_err(y, d['y']) <= 1e-4
    assert rel_err(x.grad, d['dx']) <= 1e-3
    if training and norm != 'n':
        assert rel_err(stack.npart.moving_mean, d['moving_mean']) <= 1e-5
        assert rel_err(stack.npart.moving_variance, d['moving_variance']) <= 1e-5
