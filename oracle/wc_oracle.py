"""CPU oracle for the whitening-coloring (WC) layer hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``wc_gan_b200/`` may import this module;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs do, and there only as the checker / the timed CPU baseline.

It restates, op for op, the reference's TF1.5/Keras2.0.8 graph in PyTorch-CPU
tensors (any float dtype; fp64 is the parity oracle, fp32 is the timed baseline):

* ``decorelation_normalization``  <- gan/conditional_layers.py:646-712
* ``conditional_conv11``          <- gan/conditional_layers.py:790-848
* ``factorized_conv11``           <- gan/conditional_layers.py:971-1034
* ``conv11``                      <- generator.py:49-51 (Keras Conv2D, kernel (1,1))
* ``wc_block``                    <- generator.py:13-90 (create_norm stack: _npart, _repart[_c/_u/_a])

Third-party arithmetic that is NOT vendored in /root/reference (TensorFlow 1.5.0,
Keras 2.0.8; pins from README.md:24-25) is restated from its published definition:
``tf.cholesky`` (lower factor), ``tf.matrix_triangular_solve(L, I)`` (L X = I),
``tf.svd`` of an SPD matrix, ``tf.nn.l2_normalize`` (x * rsqrt(max(sum x^2, 1e-12))),
``K.moving_average_update`` (v*m + val*(1-m); zero_debias=False assumed),
``tf.matrix_band_part(k, 0, -1)`` (upper triangle incl. diagonal of the last two dims).
Backward passes come from torch.autograd over these same ops (the reference relies on
TF autodiff, also un-vendored); ``oracle/closed_form.py`` holds the hand-derived
formulas the CUDA kernels implement and is checked against this autograd.

PARITY PINNING.  TensorFlow/Keras cannot be imported here (Python 3.12, no TF wheel, no
network), so the reference itself cannot be run.  The oracle is pinned by the reference's
own known-answer tests re-expressed in tests/test_oracle_reference_vectors.py:
gan/unit_tests.py:602-628 (Cholesky train forward, statistical 1e-3), :381-399
(class gather + bias), :521-541 (triangular, exact), :544-567 (l2-normalised
dictionary, exact) and the gather-semantics siblings.  Everything else is
**parity unpinned** by any reference test (ZCA forward, renorm, inference path, EMA
values, every gradient, ucconv/ufconv sums) and rests on this restatement alone; group
whitening (g < C) does not exist in the reference and is defined here as the reference
function applied to each contiguous channel slice.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence, Tuple

import torch

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------
# DecorelationNormalization  (gan/conditional_layers.py:606-723)
# --------------------------------------------------------------------------------------
def _get_inv_sqrt(ff: Tensor, decomposition: str, epsilon: float) -> Tuple[Tensor, Tensor]:
    """get_inv_sqrt, gan/conditional_layers.py:662-678.  Returns (sqrt, inv_sqrt)."""
    c = ff.shape[0]
    eye = torch.eye(c, dtype=ff.dtype)
    if decomposition == 'cholesky':
        sqrt = torch.linalg.cholesky(ff)                                        # :665
        inv_sqrt = torch.linalg.solve_triangular(sqrt, eye, upper=False)        # :666
        return sqrt, inv_sqrt
    if decomposition == 'zca':
        # :671  S, U, _ = svd(ff + eye*eps)   (pinned to the CPU in the reference)
        U, S, _ = torch.linalg.svd(ff + eye * epsilon, full_matrices=True)
        inv_sqrt = (U @ torch.diag(S.pow(-0.5))) @ U.t()                        # :672-673
        sqrt = (U @ torch.diag(S.pow(0.5))) @ U.t()                             # :674-675
        return sqrt, inv_sqrt
    raise AssertionError(decomposition)                                         # :615


def decorelation_normalization(x: Tensor, moving_mean: Tensor, moving_cov: Tensor, *,
                               training: bool = True, momentum: float = 0.99,
                               epsilon: float = 1e-3, decomposition: str = 'cholesky',
                               renorm: bool = False):
    """DecorelationNormalization.call, gan/conditional_layers.py:646-712.

    x (N,H,W,C) NHWC; moving_mean (C,1); moving_cov (C,C).  Returns
    (out, new_moving_mean, new_moving_cov); the moving statistics are returned
    unchanged in the test phase.  ``renorm`` reads the moving covariance as it was
    before this call's update (TF evaluates the read and the assign in one
    ``sess.run`` without a control dependency).
    """
    assert decomposition in ('cholesky', 'zca')
    n, w, h, c = x.shape                                                        # :649-650
    x_t = x.permute(3, 0, 1, 2)                                                 # :652
    x_flat = x_t.reshape(c, -1)                                                 # :655
    eye = torch.eye(c, dtype=x.dtype)
    if training:
        m = x_flat.mean(dim=1, keepdim=True)                                    # :658
    else:
        m = moving_mean                                                         # :659
    f = x_flat - m                                                              # :660
    new_mm, new_mc = moving_mean, moving_cov
    if training:
        ff_apr = (f @ f.t()) / (float(n * w * h) - 1.0)                         # :683
        with torch.no_grad():                                                   # :684-690
            new_mm = moving_mean * momentum + m.detach() * (1.0 - momentum)
            new_mc = moving_cov * momentum + ff_apr.detach() * (1.0 - momentum)
        ff_apr_shrinked = (1 - epsilon) * ff_apr + eye * epsilon                # :691
        if renorm:                                                              # :693-698
            l, l_inv = _get_inv_sqrt(ff_apr_shrinked, decomposition, epsilon)
            ff_mov = (1 - epsilon) * moving_cov + eye * epsilon
            _, l_mov_inverse = _get_inv_sqrt(ff_mov, decomposition, epsilon)
            l_ndiff = l.detach()
            inv_sqrt = (l_mov_inverse @ l_ndiff) @ l_inv
        else:
            inv_sqrt = _get_inv_sqrt(ff_apr_shrinked, decomposition, epsilon)[1]  # :700
    else:
        ff_mov = (1 - epsilon) * moving_cov + eye * epsilon                     # :703
        inv_sqrt = _get_inv_sqrt(ff_mov, decomposition, epsilon)[1]             # :704
    f_hat = inv_sqrt @ f                                                        # :707
    decorelated = f_hat.reshape(c, n, w, h).permute(1, 2, 3, 0)                 # :709-710
    return decorelated, new_mm, new_mc


def group_decorelation_normalization(x: Tensor, moving_mean: Tensor, moving_cov: Tensor, *,
                                     group_size: Optional[int] = None, **kw):
    """Block-diagonal ("group") whitening: NOT in the reference (SURVEY.md finding 1).

    Defined as ``decorelation_normalization`` applied independently to each contiguous
    channel slice of width ``group_size``.  moving_mean (C,1); moving_cov (G,g,g).
    With group_size in (None, C) this is exactly the reference layer and moving_cov
    may be given as (C,C).
    """
    c = x.shape[-1]
    g = c if group_size is None else int(group_size)
    assert c % g == 0
    G = c // g
    mc = moving_cov.reshape(G, g, g)
    outs, mms, mcs = [], [], []
    for k in range(G):
        sl = slice(k * g, (k + 1) * g)
        o, mm, mcv = decorelation_normalization(x[..., sl], moving_mean[sl], mc[k], **kw)
        outs.append(o), mms.append(mm), mcs.append(mcv)
    return (torch.cat(outs, dim=-1), torch.cat(mms, dim=0),
            torch.stack(mcs, dim=0).reshape(moving_cov.shape))


# --------------------------------------------------------------------------------------
# Coloring layers
# --------------------------------------------------------------------------------------
def conv11(x: Tensor, kernel: Tensor, bias: Optional[Tensor]) -> Tensor:
    """Keras Conv2D(kernel_size=(1,1)) as used at generator.py:49-51,57,65,75.

    kernel (1,1,Cin,Cout) (or (Cin,Cout)); bias (Cout,) or None; x NHWC.
    """
    k2 = kernel.reshape(kernel.shape[-2], kernel.shape[-1])
    out = x @ k2
    if bias is not None:
        out = out + bias
    return out


def _to_bhwc_rows(x: Tensor):
    """gan/conditional_layers.py:798-805 / :979-986: NHWC -> NCHW -> (N,C,HW) -> (N,HW,C)."""
    x = x.permute(0, 3, 1, 2)
    n, in_c, w, h = x.shape
    x = x.reshape(-1, in_c, w * h)
    x = x.permute(0, 2, 1)
    return x, w, h


def _from_rows(out: Tensor, filters: int, w: int, h: int, bias_n: Optional[Tensor]) -> Tensor:
    """gan/conditional_layers.py:828-843 / :1013-1029."""
    out = out.permute(0, 2, 1)
    out = out.reshape(-1, filters, w, h)
    if bias_n is not None:
        out = out + bias_n[:, :, None, None]
    return out.permute(0, 2, 3, 1)


def conditional_conv11(x: Tensor, cls: Tensor, kernel: Tensor, bias: Optional[Tensor],
                       triangular: bool = False) -> Tensor:
    """ConditionalConv11.call, gan/conditional_layers.py:790-848 (channels_last, no activation).

    x (N,H,W,Cin); cls (N,1) integer; kernel (K,1,1,Cin,Cout); bias (K,Cout) or None.
    """
    filters = kernel.shape[-1]
    xr, w, h = _to_bhwc_rows(x)
    c = cls.squeeze(1).long()                                                   # :809
    k = torch.triu(kernel) if triangular else kernel                            # :811-814
    k = k[c]                                                                    # :815
    k = k.squeeze(1).squeeze(1)                                                 # :818-819
    out = xr @ k                                                                # :824
    b = bias[c] if bias is not None else None                                   # :834
    return _from_rows(out, filters, w, h, b)


def l2_normalize(x: Tensor, axis: int, eps: float = 1e-12) -> Tensor:
    """tf.nn.l2_normalize (TF 1.5): x * rsqrt(maximum(reduce_sum(x*x, axis), eps))."""
    sq = (x * x).sum(dim=axis, keepdim=True)
    return x * torch.rsqrt(torch.clamp(sq, min=eps))


def factorized_conv11(x: Tensor, cls: Tensor, class_matrix: Tensor, kernel: Tensor,
                      bias: Optional[Tensor]) -> Tensor:
    """FactorizedConv11.call, gan/conditional_layers.py:971-1034.

    class_matrix (K,E); kernel (E,1,1,Cin,Cout); bias (K,Cout) or None.
    """
    e = kernel.shape[0]
    in_c, filters = kernel.shape[-2], kernel.shape[-1]
    xr, w, h = _to_bhwc_rows(x)
    c = cls.squeeze(1).long()                                                   # :990
    cls_emb = class_matrix[c]                                                   # :993
    cls_emb = l2_normalize(cls_emb, axis=1)                                     # :994
    k = kernel.reshape(e, -1)                                                   # :996
    k = cls_emb @ k                                                             # :998
    k = k.reshape(-1, in_c, filters)                                            # :1001-1005
    out = xr @ k                                                                # :1010
    b = bias[c] if bias is not None else None                                   # :1018-1020
    return _from_rows(out, filters, w, h, b)


# --------------------------------------------------------------------------------------
# create_norm stack  (generator.py:13-90)
# --------------------------------------------------------------------------------------
WC_AFTER_NORMS = ('uconv', 'cconv', 'ucconv', 'fconv', 'ufconv', 'n')


def wc_block(x: Tensor, cls: Optional[Tensor], params: dict, *, norm: str = 'd',
             after_norm: str = 'uconv', training: bool = True, momentum: float = 0.99,
             epsilon: float = 1e-3, decomposition: str = 'cholesky',
             group_size: Optional[int] = None, triangular: bool = False):
    """The stack ``create_norm(norm, after_norm, ...)(axis, name)(x)`` builds, generator.py:82-88.

    params keys (Keras weight names / shapes):
      moving_mean (C,1), moving_variance (C,C) or (G,g,g)                 [_npart]
      u_kernel (1,1,C,C), u_bias (C,)                                      [_repart / _repart_u]
      c_kernel (K,1,1,C,C), c_bias (K,C)                                   [cconv: _repart / ucconv: _repart_c]
      class_matrix (K,E), f_kernel (E,1,1,C,C)                             [fconv/ufconv: _repart_c, use_bias=False]
    Returns (y, new_moving_mean, new_moving_cov).
    """
    assert norm in ('n', 'd', 'dr')
    assert after_norm in WC_AFTER_NORMS
    mm, mc = params.get('moving_mean'), params.get('moving_variance')
    if norm == 'n':                                                             # :19-20
        xh = x
    else:                                                                       # :23-26
        xh, mm, mc = group_decorelation_normalization(
            x, mm, mc, group_size=group_size, training=training, momentum=momentum,
            epsilon=epsilon, decomposition=decomposition, renorm=(norm == 'dr'))
    if after_norm == 'n':                                                       # :79-80
        y = xh
    elif after_norm == 'uconv':                                                 # :49-51
        y = conv11(xh, params['u_kernel'], params.get('u_bias'))
    elif after_norm == 'cconv':                                                 # :41-44
        y = conditional_conv11(xh, cls, params['c_kernel'], params.get('c_bias'), triangular)
    elif after_norm == 'ucconv':                                                # :52-60
        c_ = conditional_conv11(xh, cls, params['c_kernel'], params.get('c_bias'), triangular)
        u_ = conv11(xh, params['u_kernel'], params.get('u_bias'))
        y = c_ + u_
    elif after_norm == 'fconv':                                                 # :45-48 (use_bias=False)
        y = factorized_conv11(xh, cls, params['class_matrix'], params['f_kernel'], params.get('f_bias'))
    elif after_norm == 'ufconv':                                                # :69-78
        c_ = factorized_conv11(xh, cls, params['class_matrix'], params['f_kernel'], params.get('f_bias'))
        u_ = conv11(xh, params['u_kernel'], params.get('u_bias'))
        y = c_ + u_
    return y, mm, mc


# --------------------------------------------------------------------------------------
# Keras default initialisers used by the coloring layers (Keras 2.0.8 initializers.py;
# un-vendored, restated): glorot_uniform with _compute_fans' receptive-field rule.
# --------------------------------------------------------------------------------------
def glorot_uniform_(shape: Sequence[int], generator: Optional[torch.Generator] = None,
                    dtype=torch.float32) -> Tensor:
    shape = tuple(int(s) for s in shape)
    if len(shape) == 2:
        fan_in, fan_out = shape
    else:
        rf = 1
        for s in shape[:-2]:
            rf *= s
        fan_in, fan_out = shape[-2] * rf, shape[-1] * rf
    limit = math.sqrt(6.0 / (fan_in + fan_out))
    return (torch.rand(shape, generator=generator, dtype=torch.float64) * 2 - 1).mul_(limit).to(dtype)


def make_params(c: int, *, after_norm: str, number_of_classes: int = 0, filters_emb: int = 0,
                group_size: Optional[int] = None, seed: int = 0, dtype=torch.float32,
                random_bias: bool = False, random_moving: bool = False) -> dict:
    """Synthetic parameters with the reference's shapes and default initialisers
    (BASELINE.md section 4).  random_bias / random_moving perturb the zero biases and the
    identity moving statistics so that tests exercise them."""
    gen = torch.Generator().manual_seed(seed)
    g = c if group_size is None else group_size
    G = c // g
    p = {'moving_mean': torch.zeros(c, 1, dtype=dtype),
         'moving_variance': (torch.eye(c, dtype=dtype) if G == 1
                             else torch.eye(g, dtype=dtype).repeat(G, 1, 1))}
    if random_moving:
        p['moving_mean'] = torch.randn(c, 1, generator=gen, dtype=torch.float64).mul(0.3).to(dtype)
        a = torch.randn(G, g, g, generator=gen, dtype=torch.float64) / math.sqrt(g)
        mc = a @ a.transpose(1, 2) + 0.5 * torch.eye(g, dtype=torch.float64)
        p['moving_variance'] = mc.reshape(p['moving_variance'].shape).to(dtype)

    def bias(shape):
        if random_bias:
            return torch.randn(shape, generator=gen, dtype=torch.float64).mul(0.1).to(dtype)
        return torch.zeros(shape, dtype=dtype)

    if after_norm in ('uconv', 'ucconv', 'ufconv'):
        p['u_kernel'] = glorot_uniform_((1, 1, c, c), gen, dtype)
        p['u_bias'] = bias((c,))
    if after_norm in ('cconv', 'ucconv'):
        p['c_kernel'] = glorot_uniform_((number_of_classes, 1, 1, c, c), gen, dtype)
        p['c_bias'] = bias((number_of_classes, c))
    if after_norm in ('fconv', 'ufconv'):
        p['class_matrix'] = glorot_uniform_((number_of_classes, filters_emb), gen, dtype)
        p['f_kernel'] = glorot_uniform_((filters_emb, 1, 1, c, c), gen, dtype)
    return p


def make_activation(shape: Sequence[int], dist: str = 'correlated', seed: int = 0,
                    dtype=torch.float32) -> Tensor:
    """Synthetic NHWC activations: 'iid' N(0,1); 'correlated' =
    N(1, 0.5 I + 0.5 11^T) as in gan/unit_tests.py:621-623, drawn as
    sqrt(.5) z + sqrt(.5) z_common + 1; 'relu' = max(0, correlated)."""
    gen = torch.Generator().manual_seed(seed)
    n, h, w, c = shape
    z = torch.randn(n, h, w, c, generator=gen, dtype=torch.float32)
    if dist == 'iid':
        return z.to(dtype)
    zc = torch.randn(n, h, w, 1, generator=gen, dtype=torch.float32)
    x = math.sqrt(0.5) * z + math.sqrt(0.5) * zc + 1.0
    if dist == 'relu':
        x = torch.clamp(x, min=0.0)
    return x.to(dtype)
