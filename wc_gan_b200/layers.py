"""Host-side mirror of the reference's WC-layer interface (same names, constructor arguments and
call conventions), as torch.nn.Modules over the fused CUDA op.

Reference                                   here
---------------------------------------------------------------------------------------------
DecorelationNormalization                   DecorelationNormalization   gan/conditional_layers.py:606-723
keras Conv2D(kernel_size=(1,1))             Conv11                      generator.py:49-51
ConditionalConv11                           ConditionalConv11           gan/conditional_layers.py:727-898
FactorizedConv11                            FactorizedConv11            gan/conditional_layers.py:901-1085
create_norm(...)(axis, name)(x)             create_norm(...)(axis, name)(x[, cls])   generator.py:13-90

Data contract: activations are NHWC float32 CUDA tensors (Keras channels_last, gan/train.py:7);
class labels are int32 (N,1) (generator.py:102).  Conditional layers are called with a list
``layer([x, cls])`` as in the reference.  Weight names and shapes match the Keras weights
 so a .h5 -> state_dict converter is mechanical.

A stack built by create_norm runs as ONE fused op (stats -> factor -> compose -> apply);
the standalone layers run the same op with the missing half switched off.
"""
from __future__ import annotations

import math
from typing import Callable, Optional, Union

import torch
from torch import nn

from . import _lib
from .functional import WCSpec, new_info, whitening_coloring


def _glorot_uniform_(t: torch.Tensor) -> torch.Tensor:
    """Keras 2.0.8 glorot_uniform with _compute_fans' receptive-field rule (un-vendored; restated)."""
    shape = tuple(t.shape)
    if len(shape) == 2:
        fan_in, fan_out = shape
    else:
        rf = 1
        for s in shape[:-2]:
            rf *= s
        fan_in, fan_out = shape[-2] * rf, shape[-1] * rf
    limit = math.sqrt(6.0 / (fan_in + fan_out))
    with torch.no_grad():
        return t.uniform_(-limit, limit)


def _init(t: torch.Tensor, initializer) -> torch.Tensor:
    if callable(initializer):
        with torch.no_grad():
            v = initializer(tuple(t.shape))
            t.copy_(torch.as_tensor(v, dtype=t.dtype).reshape(t.shape))
        return t
    if initializer in ('glorot_uniform', None):
        return _glorot_uniform_(t)
    if initializer == 'zeros':
        with torch.no_grad():
            return t.zero_()
    if initializer == 'ones':
        with torch.no_grad():
            return t.fill_(1.0)
    raise ValueError(f'unknown initializer {initializer!r}')


class _LazyBuild(nn.Module):
    """Keras' build(input_shape) on first call (gan/conditional_layers.py:626,763,937)."""

    def __init__(self, name: Optional[str] = None):
        super().__init__()
        self.layer_name = name
        self.built = False
        self.info = None

    def _maybe_build(self, x: torch.Tensor):
        if not self.built:
            if x.shape[-1] is None:
                raise ValueError('The channel dimension of the inputs should be defined.')
            self.build(int(x.shape[-1]), x.device)
            self.built = True
        if self.info is None or self.info.device != x.device:
            self.info = new_info(x.device)

    def check_info(self):
        """Host-side read of the device status flags (synchronises).  Raises like TF would:
        a non-PD covariance is an InvalidArgumentError in the reference (swallowed at
        gan/train.py:109-112); a label outside [0,K) is an out-of-range gather."""
        if self.info is None:
            return
        v = self.info.tolist()
        self.info.zero_()
        if v[_lib.INFO_NOT_PD]:
            raise ArithmeticError(f'{self.layer_name}: covariance not positive definite at pivot {v[_lib.INFO_NOT_PD] - 1}')
        if v[_lib.INFO_BAD_CLASS]:
            raise IndexError(f'{self.layer_name}: class label out of range for sample {v[_lib.INFO_BAD_CLASS] - 1}')
        if v[_lib.INFO_XCHG_TIMEOUT]:
            raise TimeoutError(f'{self.layer_name}: synchronised-batch exchange timed out waiting for rank '
                               f'{v[_lib.INFO_XCHG_TIMEOUT] - 1}')


def _split_inputs(inputs):
    if isinstance(inputs, (list, tuple)):
        if len(inputs) != 2:
            raise ValueError('conditional layers are called as layer([x, cls])')
        return inputs[0], inputs[1]
    raise ValueError('conditional layers are called as layer([x, cls])')


class DecorelationNormalization(_LazyBuild):
    """Batch whitening, gan/conditional_layers.py:606-723.  ``group_size`` (None = all channels,
    the reference behaviour) and ``process_group`` (synchronised-batch statistics) are additions."""

    def __init__(self, momentum=0.99, epsilon=1e-3, moving_mean_initializer='zeros', decomposition='cholesky',
                 renorm=False, moving_cov_initializer='identity', group_size=None, channels=None,
                 process_group=None, name=None, **kwargs):
        assert decomposition in ['cholesky', 'zca']                                    # :615
        super().__init__(name)
        self.supports_masking = True
        self.momentum, self.epsilon = momentum, epsilon
        self.moving_mean_initializer, self.moving_cov_initializer = moving_mean_initializer, moving_cov_initializer
        self.axis = -1                                                                 # :622
        self.renorm, self.decomposition = renorm, decomposition
        self.group_size, self.process_group = group_size, process_group
        if channels is not None:
            self.build(int(channels), None)
            self.built = True

    def build(self, dim: int, device):
        g = dim if self.group_size is None else int(self.group_size)
        if dim % g:
            raise ValueError(f'channels {dim} not divisible by group_size {g}')
        mm = torch.zeros(dim, 1, device=device)                                        # :634-638
        _init(mm, self.moving_mean_initializer)
        if g == dim:
            mc = torch.eye(dim, device=device)                                         # :639-643 name 'moving_variance'
        else:
            mc = torch.eye(g, device=device).repeat(dim // g, 1, 1)
        if callable(self.moving_cov_initializer):
            _init(mc, self.moving_cov_initializer)
        self.register_buffer('moving_mean', mm)
        self.register_buffer('moving_variance', mc)
        self.channels = dim

    @property
    def moving_cov(self):
        return self.moving_variance

    def spec(self, **over) -> WCSpec:
        d = dict(whiten=True, renorm=self.renorm, decomposition=self.decomposition, epsilon=self.epsilon,
                 momentum=self.momentum, group_size=0 if self.group_size is None else int(self.group_size))
        d.update(over)
        return WCSpec(**d)

    def forward(self, inputs, training=None):
        self._maybe_build(inputs)
        training = self.training if training is None else training
        return whitening_coloring(inputs, spec=self.spec(), training=training, moving_mean=self.moving_mean,
                                  moving_cov=self.moving_variance, group=self.process_group, info=self.info)

    def get_config(self):                                                              # :714-723
        return {'axis': self.axis, 'momentum': self.momentum, 'epsilon': self.epsilon,
                'moving_mean_initializer': self.moving_mean_initializer,
                'moving_variance_initializer': self.moving_cov_initializer}


class Conv11(_LazyBuild):
    """The unconditional coloring: keras Conv2D(kernel_size=(1,1), filters=C) at generator.py:49-51.
    Weights: kernel (1,1,Cin,Cout) glorot_uniform, bias (Cout,) zeros."""

    def __init__(self, filters, kernel_size=(1, 1), use_bias=True, kernel_initializer='glorot_uniform',
                 bias_initializer='zeros', in_channels=None, name=None, **kwargs):
        if tuple(kernel_size) != (1, 1):
            raise ValueError('only the 1x1 convolution of the WC coloring step is implemented here')
        super().__init__(name)
        self.filters, self.use_bias = filters, use_bias
        self.kernel_initializer, self.bias_initializer = kernel_initializer, bias_initializer
        if in_channels is not None:
            self.build(int(in_channels), None)
            self.built = True

    def build(self, dim: int, device):
        self.kernel = nn.Parameter(_init(torch.empty(1, 1, dim, self.filters, device=device), self.kernel_initializer))
        self.bias = nn.Parameter(_init(torch.empty(self.filters, device=device), self.bias_initializer)) if self.use_bias else None

    def weights(self):
        return dict(Wu=self.kernel.reshape(self.kernel.shape[-2], self.kernel.shape[-1]), bu=self.bias)

    def forward(self, x):
        self._maybe_build(x)
        return whitening_coloring(x, spec=WCSpec(whiten=False), training=self.training, info=self.info, **self.weights())


class ConditionalConv11(_LazyBuild):
    """gan/conditional_layers.py:727-898.  Called as layer([x, cls]); kernel (K,1,1,Cin,filters), bias (K,filters)."""

    def __init__(self, filters, number_of_classes, strides=1, data_format=None, activation=None, use_bias=True,
                 kernel_initializer='glorot_uniform', bias_initializer='zeros', kernel_regularizer=None,
                 bias_regularizer=None, activity_regularizer=None, kernel_constraint=None, bias_constraint=None,
                 triangular=False, in_channels=None, name=None, **kwargs):
        super().__init__(name)
        if strides not in (1, (1, 1)):
            raise ValueError('ConditionalConv11 only supports strides=1')
        if data_format not in (None, 'channels_last'):
            raise ValueError('only channels_last (NHWC) is supported (gan/train.py:7 asserts it)')
        self.filters, self.number_of_classes = filters, number_of_classes
        self.activation, self.use_bias, self.triangular = activation, use_bias, triangular
        self.kernel_initializer, self.bias_initializer = kernel_initializer, bias_initializer
        if in_channels is not None:
            self.build(int(in_channels), None)
            self.built = True

    def build(self, dim: int, device):
        self.input_dim = dim
        K = self.number_of_classes
        self.kernel = nn.Parameter(_init(torch.empty(K, 1, 1, dim, self.filters, device=device), self.kernel_initializer))  # :773
        self.bias = nn.Parameter(_init(torch.empty(K, self.filters, device=device), self.bias_initializer)) if self.use_bias else None

    def weights(self):
        K = self.number_of_classes
        return dict(Wc=self.kernel.reshape(K, self.kernel.shape[-2], self.kernel.shape[-1]), bc=self.bias)

    def forward(self, inputs):
        x, cls = _split_inputs(inputs)
        self._maybe_build(x)
        spec = WCSpec(whiten=False, triangular=self.triangular, number_of_classes=self.number_of_classes)
        out = whitening_coloring(x, spec=spec, training=self.training, cls=cls, info=self.info, **self.weights())
        return self.activation(out) if self.activation is not None else out                # :845-846


class FactorizedConv11(_LazyBuild):
    """gan/conditional_layers.py:901-1085: soft-assignment coloring.  class_matrix (K,E),
    kernel (E,1,1,Cin,filters), bias (K,filters)."""

    def __init__(self, filters, number_of_classes, filters_emb, strides=1, data_format=None, activation=None,
                 use_bias=True, kernel_initializer='glorot_uniform', bias_initializer='zeros',
                 kernel_regularizer=None, bias_regularizer=None, activity_regularizer=None, kernel_constraint=None,
                 bias_constraint=None, in_channels=None, name=None, **kwargs):
        super().__init__(name)
        if strides not in (1, (1, 1)):
            raise ValueError('FactorizedConv11 only supports strides=1')
        if data_format not in (None, 'channels_last'):
            raise ValueError('only channels_last (NHWC) is supported (gan/train.py:7 asserts it)')
        self.filters, self.number_of_classes, self.filters_emb = filters, number_of_classes, filters_emb
        self.activation, self.use_bias = activation, use_bias
        self.kernel_initializer, self.bias_initializer = kernel_initializer, bias_initializer
        if in_channels is not None:
            self.build(int(in_channels), None)
            self.built = True

    def build(self, dim: int, device):
        self.input_dim = dim
        K, E = self.number_of_classes, self.filters_emb
        self.class_matrix = nn.Parameter(_init(torch.empty(K, E, device=device), self.kernel_initializer))                  # :952-954
        self.kernel = nn.Parameter(_init(torch.empty(E, 1, 1, dim, self.filters, device=device), self.kernel_initializer))  # :956-960
        self.bias = nn.Parameter(_init(torch.empty(K, self.filters, device=device), self.bias_initializer)) if self.use_bias else None

    def weights(self):
        E = self.filters_emb
        return dict(class_matrix=self.class_matrix, D=self.kernel.reshape(E, self.kernel.shape[-2], self.kernel.shape[-1]),
                    bf=self.bias)

    def forward(self, inputs):
        x, cls = _split_inputs(inputs)
        self._maybe_build(x)
        spec = WCSpec(whiten=False, number_of_classes=self.number_of_classes)
        out = whitening_coloring(x, spec=spec, training=self.training, cls=cls, info=self.info, **self.weights())
        return self.activation(out) if self.activation is not None else out


class WhiteningColoring(_LazyBuild):
    """The stack create_norm returns for norm in {d, dr} x after_norm in {uconv, cconv, ucconv, fconv,
    ufconv, n}, executed as one fused op.  Sub-layers keep the reference's names:
    ``npart`` (<name>_npart), ``repart`` (<name>_repart) or ``repart_c`` / ``repart_u``
    (<name>_repart_c, <name>_repart_u; the Add <name>_repart_a has no weights)."""

    def __init__(self, norm, after_norm, number_of_classes=None, filters_emb=10, cls=None, name=None,
                 decomposition='cholesky', group_size=None, process_group=None, triangular=False,
                 uncoditional_conv_layer=None, conditional_conv_layer=None, factor_conv_layer=None,
                 momentum=0.99, epsilon=1e-3, activation=None):
        super().__init__(name)
        if activation not in (None, 'relu'):
            raise ValueError("only activation='relu' can be fused into the WC layer")
        self.activation = activation
        assert norm in ('n', 'd', 'dr')
        assert after_norm in ('uconv', 'cconv', 'ucconv', 'fconv', 'ufconv', 'n')
        self.norm, self.after_norm = norm, after_norm
        self.number_of_classes, self.filters_emb = number_of_classes, filters_emb
        self._cls = cls
        self.triangular = triangular
        self._uconv = uncoditional_conv_layer or Conv11
        self._cconv = conditional_conv_layer or ConditionalConv11
        self._fconv = factor_conv_layer or FactorizedConv11
        self.npart = None
        if norm in ('d', 'dr'):                                                        # generator.py:23-26
            self.npart = DecorelationNormalization(momentum=momentum, epsilon=epsilon, decomposition=decomposition,
                                                   renorm=(norm == 'dr'), group_size=group_size,
                                                   process_group=process_group, name=(name or '') + '_npart')

    def build(self, dim: int, device):
        nm = (self.layer_name or '') + '_repart'
        K, an = self.number_of_classes, self.after_norm
        if self.npart is not None and not self.npart.built:
            self.npart.build(dim, device)
            self.npart.built = True
        if an in ('cconv', 'ucconv', 'fconv', 'ufconv') and not K:
            raise ValueError(f"after_norm={an!r} needs number_of_classes")
        if an == 'uconv':                                                              # generator.py:49-51
            self.repart = self._uconv(kernel_size=(1, 1), filters=dim, name=nm)
        elif an == 'cconv':                                                            # :41-44
            self.repart = self._cconv(filters=dim, number_of_classes=K, name=nm, triangular=self.triangular)
        elif an == 'fconv':                                                            # :45-48
            self.repart_c = self._fconv(number_of_classes=K, name=nm + '_c', filters=dim,
                                        filters_emb=self.filters_emb, use_bias=False)
        elif an == 'ucconv':                                                           # :52-60
            self.repart_c = self._cconv(number_of_classes=K, name=nm + '_c', filters=dim, triangular=self.triangular)
            self.repart_u = self._uconv(kernel_size=(1, 1), filters=dim, name=nm + '_u')
        elif an == 'ufconv':                                                           # :69-78
            self.repart_c = self._fconv(number_of_classes=K, name=nm + '_c', filters=dim,
                                        filters_emb=self.filters_emb, use_bias=False)
            self.repart_u = self._uconv(kernel_size=(1, 1), filters=dim, name=nm + '_u')
        for m in self.children():
            if isinstance(m, _LazyBuild) and not m.built:
                m.build(dim, device)
                m.built = True

    def _labels(self, cls):
        if cls is None:
            cls = self._cls() if callable(self._cls) else self._cls
        return cls

    def forward(self, x, cls=None):
        if isinstance(x, (list, tuple)):
            x, cls = _split_inputs(x)
        self._maybe_build(x)
        w = {}
        for m in self.children():
            if m is not self.npart:
                w.update(m.weights())
        K = self.number_of_classes if self.after_norm in ('cconv', 'ucconv', 'fconv', 'ufconv') else 1
        relu = self.activation == 'relu'
        if self.npart is not None:
            spec = self.npart.spec(triangular=self.triangular, number_of_classes=K or 1, relu=relu)
            kw = dict(moving_mean=self.npart.moving_mean, moving_cov=self.npart.moving_variance,
                      group=self.npart.process_group)
        else:
            spec = WCSpec(whiten=False, triangular=self.triangular, number_of_classes=K or 1, relu=relu)
            kw = {}
        conditional = self.after_norm in ('cconv', 'ucconv', 'fconv', 'ufconv')
        return whitening_coloring(x, spec=spec, training=self.training,
                                  cls=self._labels(cls) if conditional else None, info=self.info, **kw, **w)


_OUT_OF_SCOPE = {
    'b': "norm='b' (BatchNormalization baseline)",
    'ucs': "after_norm='ucs' (CenterScale)", 'ccs': "after_norm='ccs' (ConditionalCenterScale)",
    'uccs': "after_norm='uccs'", 'ccsuconv': "after_norm='ccsuconv'",
}


def create_norm(norm, after_norm, cls=None, number_of_classes=None, filters_emb=10,
                uncoditional_conv_layer=Conv11, conditional_conv_layer=ConditionalConv11,
                factor_conv_layer=FactorizedConv11, decomposition='cholesky', group_size=None,
                process_group=None, activation=None):
    """generator.py:13-90.  Same flag vocabulary and call protocol:

        block_norm = create_norm('d', 'ufconv', cls=cls, number_of_classes=10, filters_emb=4)
        y = block_norm(axis=-1, name='Generator.0_bn1')(x)

    ``activation='relu'`` (an addition) fuses the ReLU that every block applies right after the layer
    (gan/layer_utils.py:170,178; generator.py:155) into the layer's output pass.

    ``cls`` may be a tensor, a zero-argument callable returning the current (N,1) int32 labels
    (the Keras Input tensor of generator.py:102 has no eager equivalent), or None when the
    returned stack is called as ``stack(x, cls)`` / ``stack([x, cls])``.  The diagonal /
    affine baselines ('b', 'ucs', 'ccs', 'uccs', 'ccsuconv') are outside the WC hot path
     and raise NotImplementedError.  ``decomposition``,
    ``group_size`` and ``process_group`` are additions (ZCA is commented out at generator.py:24)."""
    assert norm in ['n', 'b', 'd', 'dr']
    assert after_norm in ['ucs', 'ccs', 'uccs', 'uconv', 'fconv', 'ufconv', 'cconv', 'ucconv', 'ccsuconv', 'n']
    for flag in (norm, after_norm):
        if flag in _OUT_OF_SCOPE and flag != 'n':
            raise NotImplementedError(_OUT_OF_SCOPE[flag] + ' is not part of the whitening-coloring hot path')

    if norm == 'n' and after_norm == 'n':                                              # generator.py:19-20,79-80: identity
        def identity_norm(axis, name):
            return _Identity(activation)
        return identity_norm

    def result_norm(axis, name):
        if axis not in (-1, 3):
            raise ValueError('the WC layer normalises the last (channel) axis of NHWC tensors')
        return WhiteningColoring(norm, after_norm, number_of_classes=number_of_classes, filters_emb=filters_emb,
                                 cls=cls, name=name, decomposition=decomposition, group_size=group_size,
                                 process_group=process_group, uncoditional_conv_layer=uncoditional_conv_layer,
                                 conditional_conv_layer=conditional_conv_layer, factor_conv_layer=factor_conv_layer,
                                 activation=activation)

    return result_norm


class _Identity(nn.Module):
    """create_norm('n', 'n'): no normalisation, no coloring (the discriminator's default, discriminator.py:19)."""

    def __init__(self, activation=None):
        super().__init__()
        self.activation = activation
        self.built = True
        self.npart = None

    def build(self, dim, device):
        pass

    def check_info(self):
        pass

    def forward(self, x, cls=None):
        if isinstance(x, (list, tuple)):
            x = x[0]
        return torch.relu(x) if self.activation == 'relu' else x
