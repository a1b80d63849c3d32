"""The reference's own known-answer tests for the WC hot path (gan/unit_tests.py),
re-expressed against the CPU oracle.  These are what pins oracle/wc_oracle.py to the
reference; inputs, initialisers and thresholds are the reference's."""
import numpy as np
import torch

from oracle import wc_oracle as O


def test_decorelation():
    """gan/unit_tests.py:602-628: (10,10,10,64) from N(1, .5I + .5 11^T); mean|cov(out) - I| < 1e-3."""
    rng = np.random.RandomState(0)
    cov = 0.5 * np.eye(64) + 0.5 * np.ones((64, 64))
    x = rng.multivariate_normal(mean=np.ones(64), cov=cov, size=(10, 10, 10))
    for dtype in (torch.float32, torch.float64):
        for dec in ('cholesky', 'zca'):
            xt = torch.from_numpy(x).to(dtype)
            out, mm, mc = O.decorelation_normalization(
                xt, torch.zeros(64, 1, dtype=dtype), torch.eye(64, dtype=dtype),
                training=True, decomposition=dec, renorm=False)
            o = out.double().numpy().reshape(-1, 64)
            assert np.mean(np.abs(np.cov(o, rowvar=False) - np.eye(64))) < 1e-3
            # EMA (conditional_layers.py:684-690): mean moved 1% of the way to ~1
            assert abs(float(mm.mean()) - 0.01) < 1e-3


def test_conditional_conv11():
    """gan/unit_tests.py:381-399: ConditionalConv11 == per-sample 1x1 conv (ConditionalConv2D),
    seeded normal init, 3 classes, cls = arange(10) % 3."""
    def kernel_init(shape):
        np.random.seed(0)
        return np.random.normal(size=shape)
    kernel = torch.from_numpy(kernel_init((3, 1, 1, 10, 20))).float()
    bias = torch.from_numpy(kernel_init((3, 20))).float()
    x = torch.from_numpy(np.arange(10 * 10 * 10 * 10).reshape((10, 10, 10, 10))).float()
    cls = torch.from_numpy(np.expand_dims(np.arange(10) % 3, axis=-1)).int()
    out1 = O.conditional_conv11(x, cls, kernel, bias).numpy()
    # independent restatement of ConditionalConv2D 1x1 (conditional_layers.py:1150-1177): map over samples
    out2 = np.stack([np.einsum('hwc,cf->hwf', x[i].numpy(), kernel[int(cls[i, 0]), 0, 0].numpy())
                     + bias[int(cls[i, 0])].numpy() for i in range(10)])
    assert np.mean(np.abs(out1 - out2)) < 1e-2


def test_triangular_conv11():
    """gan/unit_tests.py:521-541: triangular keeps the upper triangle of (Cin,Cout): outputs 1 and 4."""
    a = np.empty((1, 1, 1, 2, 2))
    a[..., 0] = 1
    a[..., 1] = 2
    kernel = torch.from_numpy(a).float()
    x = torch.ones(1, 2, 2, 2)
    cls = torch.zeros(1, 1, dtype=torch.int32)
    out = O.conditional_conv11(x, cls, kernel, torch.zeros(1, 2), triangular=True).numpy()
    assert np.all(out[0, ..., 0] == 1)
    assert np.all(out[0, ..., 1] == 4)


def test_triangular_factorized_conv11():
    """gan/unit_tests.py:544-567: filters_emb=1, class_matrix rows [1],[2] -> l2-normalised to 1 -> all 2."""
    def kernel_init(shape):
        a = np.empty(shape)
        a[0] = 1
        if shape[0] != 1:
            a[1] = 2
        return a
    cm = torch.from_numpy(kernel_init((2, 1))).float()
    kernel = torch.from_numpy(kernel_init((1, 1, 1, 2, 2))).float()
    x = torch.ones(2, 2, 2, 2)
    cls = torch.arange(2, dtype=torch.int32).reshape(2, 1)
    out = O.factorized_conv11(x, cls, cm, kernel, torch.zeros(2, 2)).numpy()
    assert np.all(out[0] == 2)
    assert np.all(out[1] == 2)


def test_class_gather_reuse():
    """Gather semantics of the sibling tests (gan/unit_tests.py:497-518, 570-600): cls[2] = 0 re-uses
    class 0's kernel and bias exactly (1x1 version of those fixtures)."""
    a = np.empty((3, 1, 1, 1, 1))
    a[0], a[1], a[2] = 1, 2, 3
    b = np.empty((3, 1))
    b[0], b[1], b[2] = 1, 2, 3
    x = torch.ones(3, 2, 2, 1)
    cls = torch.tensor([[0], [1], [0]], dtype=torch.int32)
    out = O.conditional_conv11(x, cls, torch.from_numpy(a).float(), torch.from_numpy(b).float()).numpy()
    assert np.all(out[0] == 2) and np.all(out[1] == 4) and np.all(out[2] == 2)
