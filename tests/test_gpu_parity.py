"""Parity of the CUDA path (through the reference-facing modules, i.e. through the C ABI) against
the CPU oracle on identical fp32 inputs.  Tolerances are north_star's: outputs <= 1e-4,
gradients <= 1e-3 (relative to the max-abs of the tensor), class indexing bit-exact."""
import numpy as np
import pytest
import torch

from oracle import wc_oracle as O

pytestmark = pytest.mark.gpu

OUT_TOL = 1e-4
GRAD_TOL = 1e-3
DEV = 'cuda'


def _wc():
    import wc_gan_b200
    return wc_gan_b200


def rel_err(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-30))


def load_params(stack, params):
    """Copy oracle parameters (Keras names/shapes) into a built WhiteningColoring stack."""
    with torch.no_grad():
        if stack.npart is not None:
            stack.npart.moving_mean.copy_(params['moving_mean'])
            stack.npart.moving_variance.copy_(params['moving_variance'])
        an = stack.after_norm
        if an == 'uconv':
            stack.repart.kernel.copy_(params['u_kernel']); stack.repart.bias.copy_(params['u_bias'])
        elif an == 'cconv':
            stack.repart.kernel.copy_(params['c_kernel']); stack.repart.bias.copy_(params['c_bias'])
        elif an == 'ucconv':
            stack.repart_c.kernel.copy_(params['c_kernel']); stack.repart_c.bias.copy_(params['c_bias'])
            stack.repart_u.kernel.copy_(params['u_kernel']); stack.repart_u.bias.copy_(params['u_bias'])
        elif an in ('fconv', 'ufconv'):
            stack.repart_c.class_matrix.copy_(params['class_matrix']); stack.repart_c.kernel.copy_(params['f_kernel'])
            if an == 'ufconv':
                stack.repart_u.kernel.copy_(params['u_kernel']); stack.repart_u.bias.copy_(params['u_bias'])


def named_grads(stack):
    an = stack.after_norm
    out = {}
    if an == 'uconv':
        out['u_kernel'], out['u_bias'] = stack.repart.kernel.grad, stack.repart.bias.grad
    elif an == 'cconv':
        out['c_kernel'], out['c_bias'] = stack.repart.kernel.grad, stack.repart.bias.grad
    elif an == 'ucconv':
        out['c_kernel'], out['c_bias'] = stack.repart_c.kernel.grad, stack.repart_c.bias.grad
        out['u_kernel'], out['u_bias'] = stack.repart_u.kernel.grad, stack.repart_u.bias.grad
    elif an in ('fconv', 'ufconv'):
        out['class_matrix'], out['f_kernel'] = stack.repart_c.class_matrix.grad, stack.repart_c.kernel.grad
        if an == 'ufconv':
            out['u_kernel'], out['u_bias'] = stack.repart_u.kernel.grad, stack.repart_u.bias.grad
    return out


def run_case(norm, after_norm, shape, *, dec='cholesky', g=None, K=10, E=4, training=True, triangular=False,
             dist='correlated', seed=0, out_tol=OUT_TOL, grad_tol=GRAD_TOL, random_moving=True):
    W = _wc()
    n, h, w, c = shape
    x32 = O.make_activation(shape, dist, seed=seed, dtype=torch.float32)
    gen = torch.Generator().manual_seed(1)
    cls = torch.randint(0, K, (n, 1), generator=gen, dtype=torch.int32)
    params = O.make_params(c, after_norm=after_norm, number_of_classes=K, filters_emb=E, group_size=g, seed=seed + 3,
                           dtype=torch.float32, random_bias=True, random_moving=random_moving)
    dy32 = torch.randn(shape, generator=gen, dtype=torch.float32)

    # ---- oracle, fp64 on the same fp32 values
    p64 = {k: v.double() for k, v in params.items()}
    leaves = {k: v.clone().requires_grad_(True) for k, v in p64.items() if k not in ('moving_mean', 'moving_variance')}
    p64.update(leaves)
    x64 = x32.double().requires_grad_(True)
    y_ref, mm_ref, mc_ref = O.wc_block(x64, cls, p64, norm=norm, after_norm=after_norm, training=training,
                                       decomposition=dec, group_size=g, triangular=triangular)
    g_ref = torch.autograd.grad(y_ref, [x64] + list(leaves.values()), dy32.double(), allow_unused=True)
    gref = dict(zip(['x'] + list(leaves.keys()), g_ref))

    # ---- CUDA path through the reference-facing factory
    stack = W.WhiteningColoring(norm, after_norm, number_of_classes=K, filters_emb=E, name='t',
                                decomposition=dec, group_size=g, triangular=triangular)
    stack.build(c, torch.device(DEV)); stack.built = True
    stack.to(DEV)
    load_params(stack, params)
    stack.train(training)
    x = x32.to(DEV).requires_grad_(True)
    y = stack(x, cls.to(DEV))
    y.backward(dy32.to(DEV))
    torch.cuda.synchronize()
    stack.check_info()

    assert y.shape == y_ref.shape
    e = rel_err(y, y_ref)
    assert e <= out_tol, f'output rel err {e:.3e}'
    e = rel_err(x.grad, gref['x'])
    assert e <= grad_tol, f'dx rel err {e:.3e}'
    for name, gr in named_grads(stack).items():
        if gref.get(name) is None:
            continue
        e = rel_err(gr.reshape(gref[name].shape), gref[name])
        assert e <= grad_tol, f'd{name} rel err {e:.3e}'
    if training and norm != 'n':
        assert rel_err(stack.npart.moving_mean, mm_ref) <= 1e-5
        assert rel_err(stack.npart.moving_variance, mc_ref) <= 1e-5
    return stack


AFTER = ['uconv', 'cconv', 'ucconv', 'fconv', 'ufconv', 'n']


@pytest.mark.parametrize('after_norm', AFTER)
@pytest.mark.parametrize('dec', ['cholesky', 'zca'])
def test_small_all_variants(after_norm, dec):
    run_case('d', after_norm, (6, 5, 4, 32), dec=dec, K=3, E=2)


@pytest.mark.parametrize('after_norm', ['uconv', 'ucconv', 'ufconv', 'n'])
@pytest.mark.parametrize('dec', ['cholesky', 'zca'])
@pytest.mark.parametrize('g', [4, 16])
def test_group_whitening(after_norm, dec, g):
    run_case('d', after_norm, (4, 6, 6, 64), dec=dec, g=g, K=4, E=3)


@pytest.mark.parametrize('after_norm', ['uconv', 'ufconv'])
@pytest.mark.parametrize('dec', ['cholesky', 'zca'])
def test_renorm(after_norm, dec):
    run_case('dr', after_norm, (8, 4, 4, 32), dec=dec, K=3, E=2)
    run_case('dr', after_norm, (8, 4, 4, 32), dec=dec, g=8, K=3, E=2)


@pytest.mark.parametrize('after_norm', ['uconv', 'cconv', 'ufconv'])
def test_inference_phase(after_norm):
    run_case('d', after_norm, (4, 4, 4, 32), training=False, K=3, E=2)
    run_case('d', after_norm, (4, 4, 4, 32), training=False, dec='zca', g=8, K=3, E=2)


@pytest.mark.parametrize('after_norm', ['uconv', 'cconv', 'ucconv', 'fconv', 'ufconv'])
def test_coloring_only(after_norm):
    run_case('n', after_norm, (5, 3, 3, 24), K=4, E=3)


def test_triangular():
    run_case('d', 'cconv', (6, 4, 4, 16), K=3, triangular=True)
    run_case('n', 'ucconv', (6, 4, 4, 16), K=3, triangular=True)


@pytest.mark.parametrize('shape,K', [((1, 3, 3, 8), 1), ((3, 1, 1, 12), 5), ((2, 7, 5, 10), 2), ((9, 2, 3, 6), 200),
                                     ((5, 17, 3, 36), 3)])
def test_odd_shapes(shape, K):
    """C not a multiple of 4 / 32, M not a multiple of any tile, N=1, K=1, many empty classes."""
    if shape[0] * shape[1] * shape[2] > shape[3] + 1:
        run_case('d', 'ucconv', shape, K=K)
    run_case('n', 'ucconv', shape, K=K)
    run_case('n', 'ufconv', shape, K=K, E=3)


@pytest.mark.parametrize('dist', ['iid', 'correlated', 'relu'])
def test_cfg1_cholesky_uconv(dist):
    """BASELINE.json configs[0]: 64x8x8x256, DecorelationNormalization (Cholesky) + unconditional coloring."""
    run_case('d', 'uconv', (64, 8, 8, 256), dist=dist)


@pytest.mark.parametrize('after_norm', ['ucconv', 'ufconv', 'cconv', 'fconv'])
def test_cfg2_conditional(after_norm):
    """BASELINE.json configs[1]: class-conditional (K=10) and soft-assignment (E=4) coloring on 64x16x16x128."""
    run_case('d', after_norm, (64, 16, 16, 128), K=10, E=4)


@pytest.mark.parametrize('dec,g', [('zca', 16), ('zca', 32), ('zca', 64), ('cholesky', 64), ('cholesky', 256)])
def test_cfg3_group_sweep(dec, g):
    """BASELINE.json configs[2] at reduced batch (oracle-sized): 32x16x16x256, ZCA / Cholesky group sweep."""
    run_case('d', 'uconv', (32, 16, 16, 256), dec=dec, g=g)


def test_cfg3_zca_full_c():
    run_case('d', 'uconv', (16, 16, 16, 256), dec='zca', g=None)


@pytest.mark.parametrize('dist', ['iid', 'correlated', 'relu'])
def test_cfg3_full_size_vs_oracle(dist):
    """BASELINE.json's headline shape at FULL size, 256x16x16x256 (M = 65536 rows), Cholesky g = C + unconditional
    coloring, against the fp64 oracle: y, dx, d u_kernel, d u_bias and both moving statistics.  This is where the
    bf16x3 accumulation of K1 / K4 over 65536 rows has to hold the north_star tolerances."""
    run_case('d', 'uconv', (256, 16, 16, 256), dist=dist)


@pytest.mark.parametrize('after_norm,shape,K,E,dec', [('ufconv', (64, 4, 4, 128), 200, 15, 'cholesky'),
                                                       ('fconv', (48, 4, 4, 64), 40, 4, 'cholesky'),
                                                       ('ufconv', (40, 6, 6, 64), 16, 4, 'cholesky'),
                                                       ('ufconv', (256, 8, 8, 128), 200, 15, 'cholesky'),
                                                       ('ufconv', (40, 6, 6, 64), 16, 4, 'zca'),
                                                       ('fconv', (64, 4, 4, 128), 40, 4, 'zca')])
def test_factorised_many_classes(after_norm, shape, K, E, dec):
    """fconv / ufconv with K >= 3E + 4 classes (cfg5: K = 200, E = 15): the library forms T_k and every parameter gradient
    from the E + 1 basis products instead of K per-class products (wc_bwd_factorised_ok); same oracle, same tolerances.
    With ZCA the backward takes the launch-per-product path, which first rebuilds the W_k the forward did not form."""
    run_case('d', after_norm, shape, K=K, E=E, dec=dec)


@pytest.mark.parametrize('g', [64, None])
def test_cfg3_full_size_zca_vs_oracle(g):
    """ZCA whitening at the FULL 256x16x16x256 batch (BASELINE cfg3 is the ZCA group sweep): g = 64 runs the shared-memory
    Jacobi kernel, g = C = 256 the multi-CTA block Jacobi."""
    run_case('d', 'uconv', (256, 16, 16, 256), dec='zca', g=g)


@pytest.mark.parametrize('after_norm,shape,K,E,g', [('ucconv', (32, 8, 8, 256), 10, 4, None),      # 2 class groups per tile, Q in stage 0
                                                      ('ufconv', (32, 8, 8, 256), 6, 3, None),
                                                      ('ucconv', (48, 8, 8, 128), 10, 4, 64),       # two whitening groups
                                                      ('ufconv', (48, 8, 8, 128), 7, 3, 32),
                                                      ('cconv', (40, 8, 8, 128), 12, 4, None),      # no unconditional kernel
                                                      ('ucconv', (24, 4, 4, 64), 5, 2, None)])
def test_split_class_sums(after_norm, shape, K, E, g):
    """Backward chain with several classes: dWeff = sum_k W_k P_k^T is formed by groups of classes that ADD their partial
    tiles into the zeroed adjoint buffer, and the tiles of Q_k that do not fit the first round of stage 0 run in stage 1
    (wc_bwd_chain in csrc/wc_small.cu); shapes chosen to hit one / several classes per group, block-diagonal whitening
    and both placements of Q."""
    run_case('d', after_norm, shape, K=K, E=E, g=g)


def test_cfg2_full_size_conditional_vs_oracle():
    """cfg4's largest generator layer shape (256 x 32 x 32 x 128, K = 10, E = 4, ufconv) against the fp64 oracle."""
    run_case('d', 'ufconv', (256, 32, 32, 128), K=10, E=4)


def test_cfg3_full_size_properties():
    """Full 256x16x16x256: size-independent properties instead of an fp64 oracle run:
    whitened output has identity covariance / zero mean (gan/unit_tests.py:602-628 criterion), and the
    layer is exactly linear in the coloring matrix."""
    W = _wc()
    shape = (256, 16, 16, 256)
    x = O.make_activation(shape, 'correlated', seed=0).to(DEV)
    layer = W.DecorelationNormalization().to(DEV)
    layer.train()
    out = layer(x)
    o = out.reshape(-1, 256).double()
    assert float(o.mean(0).abs().max()) < 1e-4
    cov = (o.t() @ o) / (o.shape[0] - 1)
    # (1-eps) shrinkage: cov(out) = I up to eps-order terms
    assert float((cov - torch.eye(256, device=DEV, dtype=torch.float64)).abs().mean()) < 1e-3
    stack = W.WhiteningColoring('d', 'uconv', name='p').to(DEV)
    stack.build(256, torch.device(DEV)); stack.built = True
    stack.train()
    with torch.no_grad():
        stack.repart.bias.normal_()
        y1 = stack(x)
        ref = out @ stack.repart.kernel.reshape(256, 256) + stack.repart.bias
    assert rel_err(y1, ref) < 1e-4


# ---- the reference's own tests, run against the CUDA modules (gan/unit_tests.py) -------------------
def test_ref_decorelation():
    """gan/unit_tests.py:602-628."""
    W = _wc()
    rng = np.random.RandomState(0)
    cov = 0.5 * np.eye(64) + 0.5 * np.ones((64, 64))
    x = rng.multivariate_normal(mean=np.ones(64), cov=cov, size=(10, 10, 10)).astype(np.float32)
    for dec in ('cholesky', 'zca'):
        layer = W.DecorelationNormalization(renorm=False, decomposition=dec).to(DEV)
        layer.train()
        out = layer(torch.from_numpy(x).to(DEV)).detach().cpu().numpy()
        out = np.reshape(out, [-1, out.shape[-1]])
        assert np.mean(np.abs(np.cov(out, rowvar=False) - np.eye(64))) < 1e-3


def test_ref_conditional_conv11():
    """gan/unit_tests.py:381-399 (ConditionalConv11 vs per-sample 1x1 convolution)."""
    W = _wc()

    def kernel_init(shape):
        np.random.seed(0)
        return np.random.normal(size=shape)
    layer = W.ConditionalConv11(number_of_classes=3, filters=20, kernel_initializer=kernel_init,
                                bias_initializer=kernel_init)
    x = np.arange(10 * 10 * 10 * 10).reshape((10, 10, 10, 10)).astype(np.float32)
    cls = np.expand_dims(np.arange(10) % 3, axis=-1).astype(np.int32)
    layer.build(10, None); layer.built = True
    layer.to(DEV)
    out1 = layer([torch.from_numpy(x).to(DEV), torch.from_numpy(cls).to(DEV)]).detach().cpu().numpy()
    k = kernel_init((3, 1, 1, 10, 20)); b = kernel_init((3, 20))
    out2 = np.stack([np.einsum('hwc,cf->hwf', x[i].astype(np.float64), k[int(cls[i, 0]), 0, 0]) + b[int(cls[i, 0])]
                     for i in range(10)])
    assert np.mean(np.abs(out1 - out2)) < 1e-2


def test_ref_triangular_conv11():
    """gan/unit_tests.py:521-541: exact outputs 1 and 4."""
    W = _wc()

    def kernel_init(shape):
        a = np.empty(shape); a[..., 0] = 1; a[..., 1] = 2
        return a
    layer = W.ConditionalConv11(number_of_classes=1, filters=2, triangular=True, kernel_initializer=kernel_init,
                                bias_initializer='zeros')
    layer.build(2, None); layer.built = True
    layer.to(DEV)
    out = layer([torch.ones(1, 2, 2, 2, device=DEV), torch.zeros(1, 1, dtype=torch.int32, device=DEV)]).detach().cpu().numpy()
    assert np.all(out[0, ..., 0] == 1)
    assert np.all(out[0, ..., 1] == 4)


def test_ref_triangular_factorized_conv11():
    """gan/unit_tests.py:544-567: exact output 2 for both classes."""
    W = _wc()

    def kernel_init(shape):
        a = np.empty(shape); a[0] = 1
        if shape[0] != 1:
            a[1] = 2
        return a
    layer = W.FactorizedConv11(number_of_classes=2, filters=2, filters_emb=1, kernel_initializer=kernel_init,
                               bias_initializer='zeros')
    layer.build(2, None); layer.built = True
    layer.to(DEV)
    cls = torch.arange(2, dtype=torch.int32, device=DEV).reshape(2, 1)
    out = layer([torch.ones(2, 2, 2, 2, device=DEV), cls]).detach().cpu().numpy()
    assert np.all(out[0] == 2)
    assert np.all(out[1] == 2)


# ---- class indexing is bit-exact -------------------------------------------------------------------
def test_class_indexing_bit_exact():
    W = _wc()
    K, c = 7, 32
    layer = W.ConditionalConv11(number_of_classes=K, filters=c, in_channels=c, bias_initializer=lambda s: np.random.RandomState(3).normal(size=s))
    layer.to(DEV)
    gen = torch.Generator().manual_seed(5)
    base = torch.randn(1, 6, 6, c, generator=gen)
    x = base.repeat(12, 1, 1, 1).to(DEV)                       # identical inputs in every sample
    cls = torch.randint(0, K, (12, 1), generator=gen, dtype=torch.int32).to(DEV)
    y = layer([x, cls])
    for i in range(12):
        for j in range(12):
            if int(cls[i]) == int(cls[j]):
                assert torch.equal(y[i], y[j])                 # same class, same input => same bits
    perm = torch.randperm(12, generator=gen).to(DEV)
    y2 = layer([x, cls[perm]])
    assert torch.equal(y2, y[perm])                            # permuting labels permutes outputs exactly
    # a label outside [0,K) is flagged, never read out of bounds
    bad = cls.clone(); bad[3] = K + 5
    layer([x, bad]); torch.cuda.synchronize()
    with pytest.raises(IndexError):
        layer.check_info()


def test_not_pd_is_reported():
    W = _wc()
    layer = W.DecorelationNormalization(epsilon=0.0).to(DEV)
    layer.train()
    x = torch.ones(4, 2, 2, 8, device=DEV)                     # zero covariance, eps = 0 => not PD
    layer(x); torch.cuda.synchronize()
    with pytest.raises(ArithmeticError):
        layer.check_info()


def test_cpu_tensor_is_rejected():
    W = _wc()
    layer = W.DecorelationNormalization()
    with pytest.raises(RuntimeError):
        layer(torch.zeros(2, 2, 2, 4))


def test_create_norm_protocol():
    """generator.py:131-154 usage: factory -> (axis, name) -> stack(x), labels captured by closure."""
    W = _wc()
    cls = torch.randint(0, 10, (8, 1), dtype=torch.int32, device=DEV)
    block_norm = W.create_norm('d', 'ufconv', cls=lambda: cls, number_of_classes=10, filters_emb=4)
    last_norm = W.create_norm('d', 'uconv', cls=cls, number_of_classes=10)
    x = torch.randn(8, 4, 4, 64, device=DEV)
    s1 = block_norm(axis=-1, name='Generator.0_bn1')
    s2 = last_norm(axis=-1, name='Generator.BN.Final')
    y = s2(s1(x))
    assert y.shape == x.shape
    assert set(s1.state_dict().keys()) == {'npart.moving_mean', 'npart.moving_variance', 'repart_c.class_matrix',
                                           'repart_c.kernel', 'repart_u.kernel', 'repart_u.bias'}
    with pytest.raises(NotImplementedError):
        W.create_norm('b', 'ucs')


# ---- both device paths stay covered: tcgen05 (default for C in {128,256}) and FP32 FFMA ------------
@pytest.fixture
def ffma_only():
    from wc_gan_b200 import functional as F
    from wc_gan_b200._lib import lib
    old = F.USE_TENSOR_CORES
    F.USE_TENSOR_CORES = False
    lib.wc_set_tensor_core_gram(0)
    yield
    F.USE_TENSOR_CORES = old
    lib.wc_set_tensor_core_gram(1)


@pytest.mark.parametrize('shape,after_norm', [((64, 8, 8, 256), 'uconv'), ((64, 16, 16, 128), 'ucconv'),
                                              ((64, 16, 16, 128), 'ufconv')])
def test_ffma_path_on_tensor_core_shapes(ffma_only, shape, after_norm):
    run_case('d', after_norm, shape, K=10, E=4)


@pytest.mark.parametrize('shape', [(3, 5, 7, 128), (2, 16, 16, 256), (5, 4, 4, 128), (1, 1, 1, 256)])
def test_tensor_core_ragged_tiles(shape):
    """Row tiles that are partly empty (HW not a multiple of 128, tiny samples) on the tcgen05 path."""
    from wc_gan_b200 import _lib
    assert _lib.lib.wc_path_for(shape[0] * shape[1] * shape[2], shape[3], shape[3], shape[3]) & 1
    run_case('n', 'ucconv', shape, K=3)
    if shape[0] * shape[1] * shape[2] > shape[3] + 1:
        run_case('d', 'ufconv', shape, K=3, E=2)


@pytest.mark.parametrize('shape', [(3, 8, 8, 128), (5, 10, 10, 128), (37, 8, 8, 256), (150, 16, 16, 128)])
def test_tensor_core_unconditional_tiling(shape):
    """Unconditional layers view all samples as one run of rows: last tile partial (rows not a multiple of 128),
    fewer tiles than SMs, and several tiles per CTA."""
    from wc_gan_b200 import _lib
    assert _lib.lib.wc_path_for(shape[0] * shape[1] * shape[2], shape[3], shape[3], shape[3]) & 1
    run_case('d', 'uconv', shape)
    run_case('n', 'uconv', shape)
