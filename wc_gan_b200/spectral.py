"""Spectral-normalised coloring layers: the only other operators of the reference that touch
the coloring matrices (``--generator_spectral 1``, generator.py:104-113).

Reference                                   here
---------------------------------------------------------------------------------------------
max_singular_val                            max_singular_val            gan/spectral_normalized_layers.py:11-31
SNConv2D (1x1, not conv_singular)           SNConv11                    :56-109
SNConditionalConv11                         SNConditionalConv11         :181-231
SNFactorizedConv11                          SNFactorizedConv11          :234-282

The power iteration is weight-space work on C x C matrices (K or E of them): it runs as a handful of torch matmuls on the
parameters, outside the M x C hot path; the normalised kernels then enter the fused WC op exactly like plain ones, so a
stack built with these layers still costs one foreign call per direction.  As in the reference, the iterate ``u`` is a
non-trainable weight named ``largest_singular_value`` drawn from N(0, 1) and updated on every call.
"""
from __future__ import annotations

import torch
from torch import nn

from .layers import ConditionalConv11, Conv11, FactorizedConv11


def _l2n(x: torch.Tensor) -> torch.Tensor:
    """K.l2_normalize over the last two axes: x * rsqrt(max(sum x^2, 1e-12))."""
    return x * torch.rsqrt(torch.clamp((x * x).sum(dim=(-1, -2), keepdim=True), min=1e-12))


def max_singular_val(w: torch.Tensor, u: torch.Tensor, fully_differentiable: bool = False, ip: int = 1):
    """gan/spectral_normalized_layers.py:11-31.  w (..., m, n), u (..., m) -> (sigma (...), u_bar (..., m))."""
    w_ = w if fully_differentiable else w.detach()
    u_bar = u.unsqueeze(-1)
    v_bar = None
    for _ in range(ip):
        v_bar = _l2n(w_.transpose(-1, -2) @ u_bar)
        u_bar = _l2n(w_ @ v_bar)
    sigma = (u_bar.transpose(-1, -2) @ (w @ v_bar)).squeeze(-1).squeeze(-1)
    return sigma, u_bar.squeeze(-1)


class _SNMixin:
    def _sn_init(self, fully_diff_spectral=True, spectral_iterations=1, renormalize=False):
        # plain attributes, set BEFORE the base constructor: it may call build() right away (in_channels given)
        object.__setattr__(self, 'fully_diff_spectral', fully_diff_spectral)
        object.__setattr__(self, 'spectral_iterations', spectral_iterations)
        object.__setattr__(self, 'renormalize', renormalize)

    def _sn_build(self, nmat: int, rows: int, device):
        shape = (nmat * rows,) if (self.renormalize or nmat == 1) else (nmat, rows)
        self.register_buffer('largest_singular_value', torch.randn(shape, device=device))     # RandomNormal(0, 1)

    def _normalised(self, kernel: torch.Tensor, nmat: int) -> torch.Tensor:
        """kernel (nmat, Cin, Cout) -> kernel / sigma (one sigma per matrix, or one for all when renormalize)."""
        u = self.largest_singular_value.detach().clone()               # a constant of this call; the buffer is updated below
        if self.renormalize or nmat == 1:
            w = kernel.reshape(-1, kernel.shape[-1])
            sigma, u_bar = max_singular_val(w, u, self.fully_diff_spectral, self.spectral_iterations)
            out = kernel / sigma
        else:
            sigma, u_bar = max_singular_val(kernel, u, self.fully_diff_spectral, self.spectral_iterations)
            out = kernel / sigma.reshape(nmat, 1, 1)
        with torch.no_grad():
            self.largest_singular_value.copy_(u_bar.detach())          # K.update(self.u, u_bar), applied on every call
        return out


class SNConv11(Conv11, _SNMixin):
    def __init__(self, *a, fully_diff_spectral=True, spectral_iterations=1, conv_singular=False, stateful=False, **kw):
        self._sn_init(fully_diff_spectral, spectral_iterations)
        Conv11.__init__(self, *a, **kw)

    def build(self, dim, device):
        Conv11.build(self, dim, device)
        self._sn_build(1, dim, device)

    def weights(self):
        w = Conv11.weights(self)
        w['Wu'] = self._normalised(w['Wu'].unsqueeze(0), 1).squeeze(0)
        return w


class SNConditionalConv11(ConditionalConv11, _SNMixin):
    def __init__(self, *a, fully_diff_spectral=True, spectral_iterations=1, stateful=False, renormalize=False, **kw):
        self._sn_init(fully_diff_spectral, spectral_iterations, renormalize)
        ConditionalConv11.__init__(self, *a, **kw)

    def build(self, dim, device):
        ConditionalConv11.build(self, dim, device)
        self._sn_build(self.number_of_classes, dim, device)

    def weights(self):
        w = ConditionalConv11.weights(self)
        w['Wc'] = self._normalised(w['Wc'], self.number_of_classes)
        return w


class SNFactorizedConv11(FactorizedConv11, _SNMixin):
    def __init__(self, *a, fully_diff_spectral=True, spectral_iterations=1, stateful=False, renormalize=False, **kw):
        self._sn_init(fully_diff_spectral, spectral_iterations, renormalize)
        FactorizedConv11.__init__(self, *a, **kw)

    def build(self, dim, device):
        FactorizedConv11.build(self, dim, device)
        self._sn_build(self.filters_emb, dim, device)

    def weights(self):
        w = FactorizedConv11.weights(self)
        w['D'] = self._normalised(w['D'], self.filters_emb)
        return w
