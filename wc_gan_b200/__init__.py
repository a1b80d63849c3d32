"""wc_gan_b200: the whitening-coloring (WC) layer of AliaksandrSiarohin/wc-gan as hand-written
sm_100a CUDA kernels behind a C ABI (include/wc_b200.h), with the reference's layer interface.

There is no CPU implementation in this package: importing it loads (or builds) libwc_b200.so
and every layer raises on non-CUDA tensors.  The CPU restatement of the reference lives in
``oracle/`` and is test infrastructure only."""
from . import _lib
from .functional import WCSpec, whitening_coloring, new_info
from .layers import (ConditionalConv11, Conv11, DecorelationNormalization, FactorizedConv11,
                     WhiteningColoring, create_norm)

from .blocks import DCBlock, Discriminator, Generator, ResBlock, make_discriminator, make_generator
from .spectral import SNConditionalConv11, SNConv11, SNFactorizedConv11, max_singular_val
from .train import GanTrainer, discriminator_loss, generator_loss
from . import checkpoint

__all__ = ['ResBlock', 'DCBlock', 'Generator', 'Discriminator', 'make_generator', 'make_discriminator', 'SNConv11',
           'SNConditionalConv11', 'SNFactorizedConv11', 'max_singular_val', 'GanTrainer', 'generator_loss',
           'discriminator_loss', 'checkpoint', 'WCSpec', 'whitening_coloring', 'new_info', 'ConditionalConv11', 'Conv11',
           'DecorelationNormalization', 'FactorizedConv11', 'WhiteningColoring', 'create_norm']
__version__ = '0.1.0'
