"""autograd.Function over the libwc_b200 C ABI: one fused whitening-coloring op.

    y[n] = (x[n] - mean) . Weff^T . W_{cls[n]} + b_{cls[n]}

Replaces the graph that generator.py:82-88 builds out of DecorelationNormalization
(gan/conditional_layers.py:646-712) followed by Conv2D 1x1 / ConditionalConv11 (:790-848) /
FactorizedConv11 (:971-1034) and Add (generator.py:58,76).  PyTorch is plumbing here (device
memory, streams, autograd graph, torch.distributed); every FLOP runs in libwc_b200.so.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch
import torch.distributed as dist

from . import _lib, sync
from ._lib import check, lib, ptr


@dataclass(frozen=True)
class WCSpec:
    """Static description of one WC stack (what create_norm's flags select)."""
    whiten: bool = True              # norm in ('d','dr')
    renorm: bool = False             # norm == 'dr'                      conditional_layers.py:693-698
    decomposition: str = 'cholesky'  # or 'zca'                          :615
    epsilon: float = 1e-3            # :609
    momentum: float = 0.99           # :608
    group_size: int = 0              # 0 = all channels (the reference); extension otherwise
    triangular: bool = False         # ConditionalConv11(triangular=True) :811-812
    number_of_classes: int = 1


# When set to a dict, every C-ABI stage is bracketed by CUDA events on the launching stream:
#   PROFILE[stage] = [(start_event, end_event), ...]      (bench.py reads the per-stage durations)
PROFILE = None

# False forces the FP32 FFMA kernels everywhere (tests compare the two device paths).
USE_TENSOR_CORES = True


def _call(name: str, *args) -> None:
    fn = getattr(lib, name)
    if PROFILE is None:
        check(fn(*args), name)
        return
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    check(fn(*args), name)
    e.record()
    PROFILE.setdefault(name, []).append((s, e))


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f'{name} must be a CUDA tensor: wc_gan_b200 has no CPU path '
                           f'(the CPU restatement lives in oracle/ and is test-only)')
    if t.dtype != torch.float32:
        raise TypeError(f'{name} must be float32 (got {t.dtype})')


def _f32(*shape, device):
    return torch.empty(shape, dtype=torch.float32, device=device)


def _ws(nbytes: int, device):
    return torch.empty((max(int(nbytes), 16) + 3) // 4, dtype=torch.float32, device=device)


_EYES = {}


def _eye(c: int, device) -> torch.Tensor:
    key = (c, str(device))
    if key not in _EYES:
        _EYES[key] = torch.eye(c, dtype=torch.float32, device=device)
    return _EYES[key]


def new_info(device) -> torch.Tensor:
    return torch.zeros(_lib.INFO_SLOTS, dtype=torch.int32, device=device)


class WhiteningColoringFn(torch.autograd.Function):
    """forward(x, Wu, bu, Wc, bc, class_matrix, D, bf, cls, moving_mean, moving_cov, spec, training, group, info)"""

    @staticmethod
    def forward(ctx, x, Wu, bu, Wc, bc, class_matrix, D, bf, cls, moving_mean, moving_cov,
                spec: WCSpec, training: bool, group, info):
        _need_cuda(x, 'x')
        if x.dim() != 4:
            raise ValueError('x must be NHWC (N,H,W,C)')
        x = x.contiguous()
        N, H, W_, Cin = x.shape
        HW = H * W_
        dev = x.device
        st = _stream()
        K = int(spec.number_of_classes) if (Wc is not None or class_matrix is not None) else 1
        if K > 1:
            if cls is None:
                raise ValueError('conditional coloring needs cls (N,1) int32')
            cls = cls.reshape(-1)
            if cls.numel() != N:
                raise ValueError(f'cls must have one label per sample (got {cls.numel()} for N={N})')
            if cls.dtype != torch.int32:
                cls = cls.to(torch.int32)
            cls = cls.contiguous()
        else:
            cls = None
        for name, w in (('Wu', Wu), ('Wc', Wc), ('D', D)):
            if w is not None and w.shape[-2] != Cin:
                raise ValueError(f'{name} has {w.shape[-2]} input channels, x has {Cin}')
        Cout = next((w.shape[-1] for w in (Wu, Wc, D) if w is not None), Cin)
        E = class_matrix.shape[1] if class_matrix is not None else 0
        g = spec.group_size or Cin
        if spec.whiten and Cin % g:
            raise ValueError(f'channels {Cin} not divisible by group_size {g}')
        G = Cin // g
        dec = _lib.WC_DECOMP[spec.decomposition]
        world = sync.world_size(group)
        do_sync = spec.whiten and training and world > 1

        mean = Weff = Wm = R = U = lam = count_dev = None
        count = float(N * HW)
        if spec.whiten:
            Weff = _f32(G, g, g, device=dev)
            if dec == 1:
                U, lam = _f32(G, g, g, device=dev), _f32(Cin, device=dev)
            if training:
                rec = _f32(Cin + G * g * g + 1, device=dev)      # {mean, S, count}: what a shard contributes
                mean_l, S_l = rec[:Cin], rec[Cin:Cin + G * g * g]
                nb = lib.wc_stats_workspace_bytes(N * HW, Cin, g)
                ws = _ws(nb, dev)
                _call('wc_stats_fwd', ptr(x), N * HW, Cin, g, ptr(ws), nb, ptr(mean_l), ptr(S_l), st)
                if do_sync:
                    rec[-1:].fill_(count)
                    recs = sync.gather_shard_records(rec, group)
                    mean, S, count_dev = _f32(Cin, device=dev), _f32(G, g, g, device=dev), _f32(1, device=dev)
                    _call('wc_stats_merge', ptr(recs), world, Cin, g, ptr(mean), ptr(S), ptr(count_dev), st)
                else:
                    mean, S = mean_l, S_l
                Wm = _f32(G, g, g, device=dev)
                R = _f32(G, g, g, device=dev) if spec.renorm else None
            else:
                mean, S = moving_mean.reshape(-1), None
            nb = lib.wc_factor_workspace_bytes(Cin, g, dec, int(spec.renorm and training))
            ws = _ws(nb, dev)
            _call('wc_factor_fwd', ptr(mean) if training else None, ptr(S), ptr(count_dev), count, Cin, g,
                                    spec.epsilon, spec.momentum, dec, int(spec.renorm), int(training),
                                    ptr(moving_mean), ptr(moving_cov), ptr(Weff),
                                    ptr(Wm if training else Weff), ptr(R), ptr(U), ptr(lam),
                                    ptr(info), ptr(ws), nb, st)
            if not training:
                Wm = Weff

        Wu_k = Wu
        if Wu is None and Wc is None and D is None:       # after_norm == 'n': identity coloring (generator.py:79-80)
            Wu_k = _eye(Cin, dev)
        Wk = _f32(K, Cin, Cout, device=dev)
        T = _f32(K, Cin, Cout, device=dev) if spec.whiten else Wk
        b = _f32(K, Cout, device=dev) if any(t is not None for t in (bu, bc, bf)) else None
        ehat = _f32(K, E, device=dev) if class_matrix is not None else None
        _call('wc_compose_fwd', ptr(Weff), Cin, Cout, g, K, ptr(Wu_k), ptr(bu), ptr(Wc), ptr(bc),
                                 int(spec.triangular), ptr(class_matrix), ptr(D), E, ptr(bf),
                                 ptr(ehat), ptr(Wk), ptr(T), ptr(b), st)
        y = _f32(N, H, W_, Cout, device=dev)
        band = g if (spec.whiten and Wu is None and Wc is None and D is None and g < Cin) else 0
        use_tc = bool(lib.wc_path_for(N * HW, Cin, Cout, g) & 1) and USE_TENSOR_CORES
        Tp = None
        if use_tc:
            Tp = _ws(lib.wc_pack_bytes(Cin, K, 0), dev)
            _call('wc_pack_fwd', ptr(T), Cin, K, ptr(Tp), st)
        _call('wc_apply_fwd', ptr(x), N, HW, Cin, Cout, ptr(mean), ptr(T), ptr(Tp), ptr(b), ptr(cls), K, band,
                               ptr(y), ptr(info), st)

        ctx.spec, ctx.training, ctx.group, ctx.sync = spec, training, group, do_sync
        ctx.dims = (N, HW, Cin, Cout, g, K, E, count)
        ctx.has = tuple(t is not None for t in (Wu, bu, Wc, bc, class_matrix, D, bf))
        ctx.save_for_backward(x, mean, cls, Weff, Wm, R, U, lam, Wk, T, ehat, count_dev, class_matrix, D)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, mean, cls, Weff, Wm, R, U, lam, Wk, T, ehat, count_dev, class_matrix, D = ctx.saved_tensors
        spec, training = ctx.spec, ctx.training
        N, HW, Cin, Cout, g, K, E, count = ctx.dims
        has_Wu, has_bu, has_Wc, has_bc, has_cm, has_D, has_bf = ctx.has
        need = ctx.needs_input_grad
        dev = x.device
        st = _stream()
        _need_cuda(dy, 'grad_output')
        dy = dy.contiguous()
        dec = _lib.WC_DECOMP[spec.decomposition]
        G = Cin // g

        P, s = _f32(K, Cin, Cout, device=dev), _f32(K, Cout, device=dev)
        nb = lib.wc_reduce_workspace_bytes(N, HW, Cin, Cout, K)
        ws = _ws(nb, dev)
        _call('wc_reduce_bwd', ptr(x), ptr(dy), N, HW, Cin, Cout, ptr(mean), ptr(cls), K, ptr(P), ptr(s),
                                ptr(ws), nb, st)

        dWu = _f32(Cin, Cout, device=dev) if (has_Wu and need[1]) else None
        dbu = _f32(Cout, device=dev) if (has_bu and need[2]) else None
        dWc = _f32(K, Cin, Cout, device=dev) if (has_Wc and need[3]) else None
        dbc = _f32(K, Cout, device=dev) if (has_bc and need[4]) else None
        dcm = _f32(K, E, device=dev) if (has_cm and need[5]) else None
        dD = _f32(E, Cin, Cout, device=dev) if (has_D and need[6]) else None
        dbf = _f32(K, Cout, device=dev) if (has_bf and need[7]) else None
        stats_grad = spec.whiten and training and need[0]
        red = _f32(G * g * g + Cin, device=dev) if stats_grad else None
        dWeff = red[:G * g * g] if stats_grad else None
        gbar = red[G * g * g:] if stats_grad else None
        nb = lib.wc_color_bwd_workspace_bytes(Cin, Cout, g, K, E)
        ws = _ws(nb, dev)
        _call('wc_color_bwd', ptr(P), ptr(s), ptr(Weff), ptr(Wk), Cin, Cout, g, K,
                               ptr(dWu), ptr(dbu), ptr(dWc), ptr(dbc), int(spec.triangular),
                               ptr(class_matrix), ptr(ehat), ptr(D), E, ptr(dD), ptr(dcm), ptr(dbf),
                               ptr(dWeff), ptr(gbar), ptr(ws), nb, st)
        dx = None
        if need[0]:
            A = cvec = None
            if stats_grad:
                if ctx.sync:
                    sync.allreduce_adjoint(red, ctx.group)
                A, cvec = _f32(G, g, g, device=dev), _f32(Cin, device=dev)
                nb = lib.wc_factor_bwd_workspace_bytes(Cin, g, dec, int(spec.renorm))
                ws = _ws(nb, dev)
                _call('wc_factor_bwd', ptr(dWeff), ptr(gbar), ptr(Weff), ptr(Wm), ptr(R), ptr(U), ptr(lam),
                                        ptr(count_dev), count, Cin, g, spec.epsilon, dec, int(spec.renorm),
                                        ptr(A), ptr(cvec), ptr(ws), nb, st)
            dx = torch.empty_like(x)
            TAp = None
            if bool(lib.wc_path_for(N * HW, Cin, Cout, g) & 1) and USE_TENSOR_CORES:
                TAp = _ws(lib.wc_pack_bytes(Cin, K, 1), dev)
                _call('wc_pack_bwd', ptr(T), ptr(A), Cin, g, K, ptr(TAp), st)
            _call('wc_apply_bwd', ptr(x), ptr(dy), N, HW, Cin, Cout, ptr(mean), ptr(T), ptr(A), ptr(TAp), ptr(cvec), g,
                                   ptr(cls), K, ptr(dx), st)

        return (dx, dWu, dbu, dWc, dbc, dcm, dD, dbf, None, None, None, None, None, None, None)


def whitening_coloring(x, *, spec: WCSpec, training: bool, moving_mean=None, moving_cov=None,
                       Wu=None, bu=None, Wc=None, bc=None, class_matrix=None, D=None, bf=None,
                       cls=None, group=None, info=None):
    """Functional entry point.  Weight layouts (row-major, contiguous, fp32):
    Wu (Cin,Cout) | bu (Cout,) | Wc (K,Cin,Cout) | bc (K,Cout) | class_matrix (K,E) | D (E,Cin,Cout) | bf (K,Cout);
    moving_mean (C,1) or (C,), moving_cov (C,C) or (G,g,g) -- updated in place when training."""
    if info is None:
        info = new_info(x.device)
    return WhiteningColoringFn.apply(x, Wu, bu, Wc, bc, class_matrix, D, bf, cls, moving_mean, moving_cov,
                                     spec, training, group, info)
