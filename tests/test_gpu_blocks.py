"""GPU tests of the rows around the hot path: the ReLU fused into K3 against the oracle, the
generator assembled from ResBlocks with WC layers, one iteration of the reference's training schedule, the SN coloring
layers inside the fused op, and the CTA-pair variant of the apply kernels."""
import ctypes

import pytest
import torch

from oracle import wc_oracle as O
from test_gpu_parity import DEV, load_params, named_grads, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('shape,after_norm,K,E', [((8, 8, 8, 256), 'uconv', 1, 0), ((12, 6, 6, 128), 'ufconv', 4, 3),
                                                  ((5, 4, 4, 24), 'ucconv', 3, 0)])
def test_fused_relu_against_oracle(shape, after_norm, K, E):
    """create_norm(..., activation='relu') == relu(reference stack), outputs and every gradient (both device paths)."""
    import wc_gan_b200 as W
    n, h, w, c = shape
    x32 = O.make_activation(shape, 'correlated', seed=4)
    gen = torch.Generator().manual_seed(2)
    cls = torch.randint(0, max(K, 1), (n, 1), generator=gen, dtype=torch.int32)
    dy32 = torch.randn(shape, generator=gen)
    params = O.make_params(c, after_norm=after_norm, number_of_classes=K, filters_emb=E, seed=9, random_bias=True, random_moving=True)
    p64 = {k: v.double() for k, v in params.items()}
    leaves = {k: v.clone().requires_grad_(True) for k, v in p64.items() if k not in ('moving_mean', 'moving_variance')}
    p64.update(leaves)
    x64 = x32.double().requires_grad_(True)
    y_pre, _, _ = O.wc_block(x64, cls, p64, norm='d', after_norm=after_norm, training=True)
    y_ref = torch.relu(y_pre)
    stack = W.create_norm('d', after_norm, number_of_classes=K, filters_emb=E, activation='relu')(axis=-1, name='t')
    stack.build(c, torch.device(DEV)); stack.built = True
    stack.to(DEV).train()
    load_params(stack, params)
    x = x32.to(DEV).requires_grad_(True)
    y = stack(x, cls.to(DEV))
    y.backward(dy32.to(DEV))
    torch.cuda.synchronize()
    assert float(y.min()) >= 0.0
    assert rel_err(y, y_ref) <= 1e-4
    # ReLU is discontinuous in its derivative: pre-activations within the forward tolerance of zero may land on either side,
    # and one flipped entry changes a whole row of dx by O(1).  The gradients are therefore compared for the SAME active set
    # (the one the device used); that the two active sets agree except near zero is what the output check above establishes.
    mask = (y.detach().cpu() > 0).double()
    assert float(((y_ref > 0).double() - mask).abs().mean()) < 1e-3
    gref = dict(zip(['x'] + list(leaves), torch.autograd.grad(y_pre, [x64] + list(leaves.values()), dy32.double() * mask, allow_unused=True)))
    assert rel_err(x.grad, gref['x']) <= 1e-3
    for name, g in named_grads(stack).items():
        if gref.get(name) is not None:
            assert rel_err(g.reshape(gref[name].shape), gref[name]) <= 1e-3, name


def test_generator_fused_relu_equals_unfused():
    """generator.py:93-163 with the cifar10_resnet_sn_cond_sa settings at reduced width: fusing the ReLUs changes nothing."""
    import wc_gan_b200 as W
    kw = dict(block_sizes=(128, 128), resamples=('UP', 'UP'), first_block_shape=(4, 4, 128), number_of_classes=10,
              block_norm='d', block_after_norm='ufconv', filters_emb=4, last_norm='d', last_after_norm='uconv')
    torch.manual_seed(0)
    torch.backends.cudnn.allow_tf32 = False            # the convolutions are not under test: keep them out of the error budget
    torch.backends.cuda.matmul.allow_tf32 = False
    G1 = W.make_generator(fuse_relu=True, **kw).to(DEV).train()
    G2 = W.make_generator(fuse_relu=False, **kw).to(DEV).train()
    G2.load_state_dict(G1.state_dict())
    # 128 samples: the first WC layer then sees 2048 rows for 128 channels (a batch covariance of 256 rows would be so
    # ill-conditioned that last-bit differences of the statistics are amplified through five whitening layers)
    z = torch.randn(128, 128, device=DEV)
    cls = torch.randint(0, 10, (128, 1), dtype=torch.int32, device=DEV)
    y1, y2 = G1(z, cls), G2(z, cls)
    assert tuple(y1.shape) == (128, 16, 16, 3) and float(y1.abs().max()) <= 1.0
    assert rel_err(y1, y2) <= 1e-3
    g1 = torch.autograd.grad(y1.square().sum(), list(G1.parameters()))
    g2 = torch.autograd.grad(y2.square().sum(), list(G2.parameters()))
    # (the statistics kernels reduce in a run-dependent order, so the two models' pre-activations differ in the last bits and
    #  a few ReLU entries sit on different sides of zero: compare in the Frobenius norm, not entry by entry)
    #  a bias in front of a whitening layer has a mathematically zero gradient -- what is computed there is rounding noise --
    #  so errors are measured against the largest gradient of the model)
    scale = max(float(b.norm()) for b in g2)
    for a, b in zip(g1, g2):
        assert float((a - b).norm()) <= 2e-2 * max(float(b.norm()), 1e-2 * scale)
    for s in G1.wc_layers():
        s.check_info()


def test_one_training_iteration_runs():
    """gan/train.py:78-99 schedule on synthetic 32x32 data (projection discriminator with spectral norm, hinge loss)."""
    import wc_gan_b200 as W
    torch.manual_seed(0)
    G = W.make_generator(number_of_classes=10, block_norm='d', block_after_norm='ufconv', filters_emb=4, last_norm='d',
                         last_after_norm='uconv').to(DEV)
    D = W.make_discriminator((32, 32, 3), block_sizes=(64, 64, 64, 64), number_of_classes=10, type='PROJECTIVE', spectral=True,
                             sum_pool=True).to(DEV)
    tr = W.GanTrainer(G, D, batch_size=8, training_ratio=2, generator_batch_multiple=2, total_generator_iters=100.0, device=DEV)
    before = [p.detach().clone() for p in G.parameters()]
    mv0 = G.wc_layers()[0].npart.moving_variance.clone()
    d, g = tr.train_one_step()
    torch.cuda.synchronize()
    assert torch.isfinite(d) and torch.isfinite(g)
    assert any(not torch.equal(a, b) for a, b in zip(before, G.parameters()))
    assert not torch.equal(mv0, G.wc_layers()[0].npart.moving_variance)     # D steps run G in training phase (gan/gan.py:254-255)
    for s in G.wc_layers():
        s.check_info()


def test_sn_coloring_inside_the_fused_op():
    """--generator_spectral 1 (generator.py:104-113): the stack with SN layers == the plain stack fed kernel / sigma."""
    import wc_gan_b200 as W
    from functools import partial
    torch.manual_seed(1)
    kw = dict(number_of_classes=4, filters_emb=3)
    sn = W.create_norm('d', 'ufconv', uncoditional_conv_layer=partial(W.SNConv11, spectral_iterations=3),
                       factor_conv_layer=partial(W.SNFactorizedConv11, spectral_iterations=3), **kw)(axis=-1, name='s')
    pl = W.create_norm('d', 'ufconv', **kw)(axis=-1, name='p')
    for s in (sn, pl):
        s.build(64, torch.device(DEV)); s.built = True
        s.to(DEV).train()
    with torch.no_grad():
        u0 = {k: v.clone() for k, v in sn.named_buffers() if 'largest_singular_value' in k}
        wsn = {}
        for m in (sn.repart_u, sn.repart_c):
            wsn.update(m.weights())
        for k, v in u0.items():                                   # weights() advanced the power iteration: rewind
            dict(sn.named_buffers())[k].copy_(v)
        pl.repart_u.kernel.copy_(wsn['Wu'].reshape(1, 1, 64, 64)); pl.repart_u.bias.copy_(sn.repart_u.bias)
        pl.repart_c.kernel.copy_(wsn['D'].reshape(3, 1, 1, 64, 64)); pl.repart_c.class_matrix.copy_(sn.repart_c.class_matrix)
    x = torch.randn(6, 5, 5, 64, device=DEV)
    cls = torch.randint(0, 4, (6, 1), dtype=torch.int32, device=DEV)
    assert rel_err(sn(x, cls), pl(x, cls)) <= 1e-5
    sn(x, cls).sum().backward()
    assert sn.repart_u.kernel.grad is not None and torch.isfinite(sn.repart_c.kernel.grad).all()


@pytest.fixture
def pair_apply():
    from wc_gan_b200._lib import lib
    lib.wc_debug_set.argtypes = [ctypes.c_int]
    lib.wc_debug_set(2097152)
    yield
    lib.wc_debug_set(0)


@pytest.mark.parametrize('shape,after_norm', [((64, 8, 8, 256), 'uconv'), ((64, 16, 16, 128), 'ufconv'), ((3, 5, 7, 128), 'ucconv'),
                                              ((37, 8, 8, 256), 'uconv')])
def test_cta_pair_apply_kernels(pair_apply, shape, after_norm):
    """tc_apply2_kernel (cta_group::2): same parity bar as the default kernel, including ragged 256-row pair tiles."""
    from test_gpu_parity import run_case
    m = shape[0] * shape[1] * shape[2]
    run_case('d' if m > 4 * shape[3] else 'n', after_norm, shape, K=10, E=4)      # whitening needs many more rows than channels
