"""The callers of the WC layer: the reference's ResNet / DCGAN blocks and its generator / discriminator, in PyTorch.

Reference                                   here
---------------------------------------------------------------------------------------------
uniform_init / he_init / glorot_init        same names                  gan/layer_utils.py:119-131
resblock(x, ...)                            ResBlock(in_filters, ...)   gan/layer_utils.py:134-188
dcblock(x, ...)                             DCBlock(in_filters, ...)    gan/layer_utils.py:191-215
make_generator(...)                         make_generator(...)         generator.py:93-163
make_discriminator(...)                     make_discriminator(...)     discriminator.py:15-85

Scope: these modules exist so that the WC layer can be measured where it runs -- inside
every generator block -- and so that the ReLU that follows every WC layer is fused into the layer's output pass
(``create_norm(..., activation='relu')``) instead of costing a read and a write of the activation.  The convolutions,
the dense layers and the pooling are plain PyTorch / cuDNN calls on channels_last views: they are not part of the hot
path this repository rewrites and nothing here tunes them.

Data contract: activations are NHWC float32 CUDA tensors between the layers (Keras channels_last, gan/train.py:7);
a convolution sees the NCHW *view* of the same memory (torch.channels_last), so no relayout copy is made.
"""
from __future__ import annotations

import math
from functools import partial
from typing import Callable, Optional, Sequence

import torch
import torch.nn.functional as F
from torch import nn

from .layers import ConditionalConv11, Conv11, FactorizedConv11, create_norm


def uniform_init_(t: torch.Tensor, constant: float = 4.0) -> torch.Tensor:
    """gan/layer_utils.py:119-129 for a Keras-shaped weight: (kh, kw, Cin, Cout) or (in, out)."""
    shape = tuple(t.shape)
    if len(shape) >= 4:
        stdev = math.sqrt(constant / ((shape[1] ** 2) * (shape[-1] + shape[-2])))
    else:
        stdev = math.sqrt(constant / (shape[0] + shape[1]))
    with torch.no_grad():
        return t.uniform_(-stdev * math.sqrt(3), stdev * math.sqrt(3))


he_init_ = partial(uniform_init_, constant=4.0)
glorot_init_ = partial(uniform_init_, constant=2.0)


class Conv2D(nn.Module):
    """keras.layers.Conv2D on NHWC tensors: kernel (kh, kw, Cin, Cout) as in Keras, 'same' padding, optional stride."""

    def __init__(self, in_filters: int, filters: int, kernel_size=(3, 3), strides=(1, 1), use_bias: bool = True,
                 kernel_initializer: Callable = he_init_, name: Optional[str] = None, transpose: bool = False):
        super().__init__()
        kh, kw = kernel_size
        self.kernel = nn.Parameter(kernel_initializer(torch.empty(kh, kw, in_filters, filters)))
        self.bias = nn.Parameter(torch.zeros(filters)) if use_bias else None
        self.strides, self.layer_name, self.transpose = tuple(strides), name, transpose

    def weight_oihw(self) -> torch.Tensor:
        return self.kernel.permute(3, 2, 0, 1)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        kh, kw = self.kernel.shape[:2]
        xin = x.permute(0, 3, 1, 2)                      # NCHW view of NHWC memory == channels_last
        if self.transpose:                               # keras Deconv2D(padding='same'): output = input * stride
            s = self.strides
            pad = ((kh - s[0] + 1) // 2, (kw - s[1] + 1) // 2)
            opad = (2 * pad[0] + s[0] - kh, 2 * pad[1] + s[1] - kw)
            y = F.conv_transpose2d(xin, self.kernel.permute(2, 3, 0, 1), self.bias, stride=s, padding=pad, output_padding=opad)
        elif self.strides == (1, 1):
            y = F.conv2d(xin, self.weight_oihw(), self.bias, padding=(kh // 2, kw // 2))
        else:                                            # TF 'same' with stride: pad so that out = ceil(in / stride)
            H, W = xin.shape[2:]
            ph = max((math.ceil(H / self.strides[0]) - 1) * self.strides[0] + kh - H, 0)
            pw = max((math.ceil(W / self.strides[1]) - 1) * self.strides[1] + kw - W, 0)
            xin = F.pad(xin, (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2))
            y = F.conv2d(xin, self.weight_oihw(), self.bias, stride=self.strides)
        return y.permute(0, 2, 3, 1).contiguous()        # back to NHWC memory (a no-op when cuDNN kept channels_last)


def _built(layer, channels: int):
    """Keras builds lazily on the first call; here the channel count is known, and parameters must exist before an
    optimiser is created over model.parameters()."""
    if layer is not None and hasattr(layer, 'build') and not getattr(layer, 'built', True):
        layer.build(int(channels), None)
        layer.built = True
    return layer


def _upsample2(x: torch.Tensor) -> torch.Tensor:
    """UpSampling2D(size=(2,2)): nearest neighbour, NHWC."""
    n, h, w, c = x.shape
    return x.reshape(n, h, 1, w, 1, c).expand(n, h, 2, w, 2, c).reshape(n, 2 * h, 2 * w, c)


def _avgpool2(x: torch.Tensor) -> torch.Tensor:
    """AveragePooling2D(pool_size=(2,2)), NHWC."""
    n, h, w, c = x.shape
    return x.reshape(n, h // 2, 2, w // 2, 2, c).mean(dim=(2, 4))


class ResBlock(nn.Module):
    """gan/layer_utils.py:134-188.  ``norm`` is what create_norm returns (called as ``norm(axis, name)``) or None;
    ``norm_relu`` is the same factory with the following ReLU fused (built by make_generator); when only ``norm`` is given
    the ReLU runs as a separate pass, exactly like the reference graph."""

    def __init__(self, in_filters: int, kernel_size, resample: str, nfilters: int, name: str, norm=None, norm_relu=None,
                 is_first: bool = False, conv_layer=Conv2D):
        super().__init__()
        assert resample in ('UP', 'SAME', 'DOWN')
        self.resample, self.is_first = resample, is_first
        self.shortcut = None
        if not (resample == 'SAME' and in_filters == nfilters):
            self.shortcut = conv_layer(in_filters, nfilters, kernel_size=(1, 1), kernel_initializer=he_init_, name=name + 'shortcut')
        self.fused = norm_relu is not None
        mk = norm_relu if norm_relu is not None else norm
        self.bn1 = _built(mk(axis=-1, name=name + '_bn1'), in_filters) if (mk is not None and not is_first) else None
        self.bn2 = _built(mk(axis=-1, name=name + '_bn2'), nfilters) if mk is not None else None
        self.conv1 = conv_layer(in_filters, nfilters, kernel_size=kernel_size, kernel_initializer=he_init_, name=name + '_conv1')
        self.conv2 = conv_layer(nfilters, nfilters, kernel_size=kernel_size, kernel_initializer=he_init_, name=name + '_conv2')

    def _resample(self, x):
        return _upsample2(x) if self.resample == 'UP' else (_avgpool2(x) if self.resample == 'DOWN' else x)

    def _norm_relu(self, layer, x, cls):
        if layer is None:
            return torch.relu(x)
        y = layer(x, cls)
        return y if self.fused else torch.relu(y)

    def forward(self, x: torch.Tensor, cls: Optional[torch.Tensor] = None) -> torch.Tensor:
        sc = (lambda t: t) if self.shortcut is None else self.shortcut
        shortcut = sc(self._resample(x)) if self.is_first else self._resample(sc(x))
        h = x
        if not self.is_first:
            h = self._norm_relu(self.bn1, h, cls)
        if self.resample == 'UP':
            h = _upsample2(h)
        h = self.conv1(h)
        h = self._norm_relu(self.bn2, h, cls)
        h = self.conv2(h)
        if self.resample == 'DOWN':
            h = _avgpool2(h)
        return shortcut + h


class DCBlock(nn.Module):
    """gan/layer_utils.py:191-215 (DCGAN block: norm -> (Leaky)ReLU -> (de)convolution)."""

    def __init__(self, in_filters: int, kernel_size, resample: str, nfilters: int, name: str, norm=None, norm_relu=None,
                 is_first: bool = False, conv_layer=Conv2D):
        super().__init__()
        assert resample in ('UP', 'SAME', 'DOWN')
        self.resample, self.is_first = resample, is_first
        self.fused = norm_relu is not None and resample == 'UP'          # only the UP branch uses a plain ReLU
        mk = norm_relu if self.fused else norm
        self.bn = _built(mk(axis=-1, name=name + '.bn'), in_filters) if (mk is not None and not (is_first and resample != 'UP')) else None
        strides = (1, 1) if resample == 'SAME' else (2, 2)
        self.conv = Conv2D(in_filters, nfilters, kernel_size=kernel_size, strides=strides, name=name + '.conv',
                           transpose=(resample == 'UP'))                    # the generator's dcblock uses Deconv2D (generator.py:149)

    def forward(self, x, cls=None):
        h = x
        if self.resample == 'UP':
            h = self.bn(h, cls) if self.bn is not None else h
            h = h if self.fused else torch.relu(h)
        elif not self.is_first:
            h = self.bn(h, cls) if self.bn is not None else h
            h = F.leaky_relu(h, 0.3)                                         # keras LeakyReLU default alpha
        return self.conv(h)


class Dense(nn.Module):
    def __init__(self, in_units: int, units: int, use_bias: bool = True, kernel_initializer=glorot_init_):
        super().__init__()
        self.kernel = nn.Parameter(kernel_initializer(torch.empty(in_units, units)))
        self.bias = nn.Parameter(torch.zeros(units)) if use_bias else None

    def forward(self, x):
        y = x @ self.kernel
        return y + self.bias if self.bias is not None else y


class Generator(nn.Module):
    """generator.py:93-163.  forward(noise (N, Z), cls (N, 1) int32) -> images (N, H, W, output_channels) in [-1, 1]."""

    def __init__(self, input_noise_shape=(128,), output_channels=3, block_sizes=(128, 128, 128), resamples=('UP', 'UP', 'UP'),
                 first_block_shape=(4, 4, 128), number_of_classes=10, concat_cls=False, block_norm='d', block_after_norm='ufconv',
                 filters_emb=10, last_norm='d', last_after_norm='uconv', arch='res', fuse_relu=True, decomposition='cholesky',
                 group_size=None, process_group=None, uncoditional_conv_layer=Conv11, conditional_conv_layer=ConditionalConv11,
                 factor_conv_layer=FactorizedConv11):
        super().__init__()
        assert arch in ('res', 'dcgan')
        self.first_block_shape = tuple(first_block_shape)
        z = input_noise_shape[0]
        self.emb = None
        if concat_cls:
            self.emb = nn.Embedding(number_of_classes, first_block_shape[-1])
            z += first_block_shape[-1]
        self.dense = Dense(z, int(math.prod(first_block_shape)), kernel_initializer=glorot_init_)
        common = dict(number_of_classes=number_of_classes, filters_emb=filters_emb, decomposition=decomposition,
                      group_size=group_size, process_group=process_group, uncoditional_conv_layer=uncoditional_conv_layer,
                      conditional_conv_layer=conditional_conv_layer, factor_conv_layer=factor_conv_layer)
        bn = create_norm(block_norm, block_after_norm, **common)
        bn_relu = create_norm(block_norm, block_after_norm, activation='relu', **common) if fuse_relu else None
        ln = create_norm(last_norm, last_after_norm, activation='relu' if fuse_relu else None, **common)
        self.fuse_relu = fuse_relu
        blocks, cin = [], first_block_shape[-1]
        for i, (bs, rs) in enumerate(zip(block_sizes, resamples)):
            bs = int(bs)
            if arch == 'res':
                blocks.append(ResBlock(cin, (3, 3), rs, bs, 'Generator.' + str(i), norm=bn, norm_relu=bn_relu))
            else:
                blocks.append(DCBlock(cin, (4, 4), rs, bs, 'Generator.' + str(i), norm=bn, norm_relu=bn_relu))
            cin = bs
        self.blocks = nn.ModuleList(blocks)
        self.last_norm = _built(ln(axis=-1, name='Generator.BN.Final'), cin)
        self.final = Conv2D(cin, output_channels, kernel_size=(3, 3), kernel_initializer=glorot_init_, name='Generator.Final')

    def wc_layers(self):
        from .layers import WhiteningColoring
        return [m for m in self.modules() if isinstance(m, WhiteningColoring)]

    def forward(self, noise: torch.Tensor, cls: Optional[torch.Tensor] = None) -> torch.Tensor:
        y = noise
        if self.emb is not None:
            y = torch.cat([self.emb(cls.reshape(-1).long()), y], dim=-1)
        y = self.dense(y).reshape(-1, *self.first_block_shape)
        for b in self.blocks:
            y = b(y, cls)
        y = self.last_norm(y, cls)
        if not self.fuse_relu:
            y = torch.relu(y)
        return torch.tanh(self.final(y))


class Discriminator(nn.Module):
    """discriminator.py:15-85.  type in {None, 'AC_GAN', 'PROJECTIVE'}; spectral=True wraps every convolution, dense and
    embedding weight in torch's spectral_norm parametrisation (gan/spectral_normalized_layers.py:56-179 for those layers:
    one power iteration per call, sigma = u^T W v)."""

    def __init__(self, input_image_shape, block_sizes=(128, 128, 128, 128), resamples=('DOWN', 'DOWN', 'SAME', 'SAME'),
                 number_of_classes=10, type='PROJECTIVE', norm='n', after_norm='n', spectral=False, sum_pool=False, dropout=0.0,
                 arch='res', filters_emb=10):
        super().__init__()
        assert arch in ('res', 'dcgan') and len(block_sizes) == len(resamples)
        self.type, self.arch, self.sum_pool = type, arch, sum_pool
        nl = create_norm(norm, after_norm, number_of_classes=number_of_classes, filters_emb=filters_emb)
        nl = None if (norm == 'n' and after_norm == 'n') else nl
        blocks, cin = [], input_image_shape[-1]
        h = input_image_shape[0]
        for i, (bs, rs) in enumerate(zip(block_sizes, resamples)):
            bs = int(bs)
            if arch == 'res':
                blocks.append(ResBlock(cin, (3, 3), rs, bs, 'Discriminator.' + str(i), norm=nl, is_first=(i == 0)))
            else:
                blocks.append(DCBlock(cin, (3, 3) if rs == 'SAME' else (4, 4), rs, bs, 'Discriminator.' + str(i), norm=nl, is_first=(i == 0)))
            cin = bs
            h = h // 2 if rs == 'DOWN' else h
        self.blocks = nn.ModuleList(blocks)
        feat = cin if arch == 'res' else cin * h * h
        self.dropout = nn.Dropout(dropout) if dropout else None
        self.out = Dense(feat, 1, kernel_initializer=glorot_init_)
        self.cls_out = Dense(feat, number_of_classes, kernel_initializer=glorot_init_) if type == 'AC_GAN' else None
        self.emb = nn.Embedding(number_of_classes, cin) if type == 'PROJECTIVE' else None
        if spectral:
            _apply_spectral_norm(self)

    def forward(self, x: torch.Tensor, cls: Optional[torch.Tensor] = None):
        y = x
        for b in self.blocks:
            y = b(y, cls)
        y = torch.relu(y) if self.arch == 'res' else F.leaky_relu(y, 0.3)
        if self.arch == 'res':
            y = y.sum(dim=(1, 2)) if self.sum_pool else y.mean(dim=(1, 2))
        else:
            y = y.reshape(y.shape[0], -1)
        if self.dropout is not None:
            y = self.dropout(y)
        if self.type == 'AC_GAN':
            return self.out(y), self.cls_out(y)
        if self.type == 'PROJECTIVE':
            phi = (self.emb(cls.reshape(-1).long()) * y).sum(dim=1, keepdim=True)
            return phi + self.out(y)
        return self.out(y)


def _apply_spectral_norm(model: nn.Module) -> None:
    from torch.nn.utils.parametrizations import spectral_norm

    # Keras reshapes a conv kernel (kh, kw, Cin, Cout) to (kh*kw*Cin, Cout) before the power iteration
    # (gan/spectral_normalized_layers.py:91-92); torch's spectral_norm flattens everything but `dim`: the same matrix, transposed
    for m in list(model.modules()):
        if isinstance(m, Conv2D):
            spectral_norm(m, name='kernel', dim=3)
        elif isinstance(m, Dense):
            spectral_norm(m, name='kernel', dim=1)
        elif isinstance(m, nn.Embedding):
            spectral_norm(m, name='weight', dim=1)


def make_generator(input_noise_shape=(128,), output_channels=3, input_cls_shape=(1,), block_sizes=(128, 128, 128),
                   resamples=('UP', 'UP', 'UP'), first_block_shape=(4, 4, 128), number_of_classes=10, concat_cls=False,
                   block_norm='d', block_after_norm='ufconv', filters_emb=10, last_norm='d', last_after_norm='uconv',
                   gan_type=None, arch='res', spectral=False, fully_diff_spectral=False, spectral_iterations=1,
                   conv_singular=True, **extensions) -> Generator:
    """generator.py:93-163, same arguments (the Keras Input shapes are accepted and ignored; ``spectral`` selects the
    SN coloring layers of wc_gan_b200.spectral).  ``extensions``: fuse_relu, decomposition, group_size, process_group."""
    layers = {}
    if spectral:
        from .spectral import SNConditionalConv11, SNConv11, SNFactorizedConv11
        kw = dict(fully_diff_spectral=fully_diff_spectral, spectral_iterations=spectral_iterations)
        layers = dict(uncoditional_conv_layer=partial(SNConv11, **kw), conditional_conv_layer=partial(SNConditionalConv11, **kw),
                      factor_conv_layer=partial(SNFactorizedConv11, **kw))
    return Generator(input_noise_shape=input_noise_shape, output_channels=output_channels, block_sizes=block_sizes,
                     resamples=resamples, first_block_shape=first_block_shape, number_of_classes=number_of_classes,
                     concat_cls=concat_cls, block_norm=block_norm, block_after_norm=block_after_norm, filters_emb=filters_emb,
                     last_norm=last_norm, last_after_norm=last_after_norm, arch=arch, **layers, **extensions)


def make_discriminator(input_image_shape, input_cls_shape=(1,), block_sizes=(128, 128, 128, 128),
                       resamples=('DOWN', 'DOWN', 'SAME', 'SAME'), number_of_classes=10, type='AC_GAN', norm='n', after_norm='n',
                       spectral=False, fully_diff_spectral=False, spectral_iterations=1, conv_singular=True, sum_pool=False,
                       dropout=False, arch='res', filters_emb=10) -> Discriminator:
    """discriminator.py:15-85, same arguments."""
    return Discriminator(input_image_shape, block_sizes=block_sizes, resamples=resamples, number_of_classes=number_of_classes,
                         type=type, norm=norm, after_norm=after_norm, spectral=bool(spectral), sum_pool=bool(sum_pool),
                         dropout=float(dropout or 0.0), arch=arch, filters_emb=filters_emb)
