"""autograd.Function over the libwc_b200 C ABI: one fused whitening-coloring op.

    y[n] = (x[n] - mean) . Weff^T . W_{cls[n]} + b_{cls[n]}

Replaces the graph that generator.py:82-88 builds out of DecorelationNormalization
(gan/conditional_layers.py:646-712) followed by Conv2D 1x1 / ConditionalConv11 (:790-848) /
FactorizedConv11 (:971-1034) and Add (generator.py:58,76).  PyTorch is plumbing here (device
memory, streams, autograd graph, torch.distributed); every FLOP runs in libwc_b200.so.

Each direction is ONE foreign call (wc_layer_fwd / wc_layer_bwd, include/wc_b200.h "whole-layer calls"):
the host fills a WcLayerCall, allocates the output, the `saved` block (what backward needs) and takes a
cached scratch workspace; the library sequences its kernels on the current stream.  With a process group
each direction is split in two phases around the collective (wc_gan_b200/sync.py).
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass
from typing import Optional

import torch
import torch.distributed as dist

from . import _lib, sync
from ._lib import WcLayerCall, WcSavedLayout, check, lib, ptr


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
    relu: bool = False               # fuse the Activation('relu') that follows the layer (gan/layer_utils.py:170,178)


# When set to a dict, every C-ABI call is bracketed by CUDA events on the launching stream:
#   PROFILE[name] = [(start_event, end_event), ...]
PROFILE = None

# False forces the FP32 FFMA kernels everywhere (tests compare the two device paths).
USE_TENSOR_CORES = True

PHASE_STATS, PHASE_APPLY = 1, 2      # forward  (WC_PHASE_*)
PHASE_REDUCE, PHASE_DX = 1, 2        # backward


NVTX = bool(os.environ.get('WC_B200_NVTX'))     # NVTX ranges around the foreign calls (nsys / ncu --nvtx timelines)


def _call(name: str, *args) -> None:
    fn = getattr(lib, name)
    if NVTX:
        torch.cuda.nvtx.range_push(name)
        try:
            check(fn(*args), name)
        finally:
            torch.cuda.nvtx.range_pop()
        return
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


# largest whitening group the factorisation kernels take (wc_factor_fwd: one group matrix per CTA / 8-CTA cluster)
MAX_GROUP = 256


def _operand(t, name, dev, shape, *, numel=None, inplace=False):
    """Validate one weight / buffer before its data_ptr() goes to the C ABI.  Returns a contiguous tensor (in-place
    buffers must already be contiguous: the kernels update them)."""
    if t is None:
        return None
    _need_cuda(t, name)
    if t.dtype != torch.float32:
        raise TypeError(f'{name} must be float32 (got {t.dtype})')
    if t.device != dev:
        raise ValueError(f'{name} is on {t.device}, x on {dev}: every operand must live on the same CUDA device')
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise ValueError(f'{name} has shape {tuple(t.shape)}, expected {tuple(shape)}')
    if numel is not None and t.numel() != numel:
        raise ValueError(f'{name} has {t.numel()} elements, expected {numel}')
    if not t.is_contiguous():
        if inplace:
            raise ValueError(f'{name} is updated in place and must be contiguous')
        t = t.contiguous()
    return t


def _f32(*shape, device):
    return torch.empty(shape, dtype=torch.float32, device=device)


def new_info(device) -> torch.Tensor:
    return torch.zeros(_lib.INFO_SLOTS, dtype=torch.int32, device=device)


# scratch workspaces: one per (device, stream, direction), grown on demand.  Work on a stream is ordered, so a
# workspace is never in use by two calls at once.
_WORKSPACES = {}


def _workspace(nbytes: int, device, stream: int, direction: int) -> torch.Tensor:
    key = (str(device), stream, direction)
    ws = _WORKSPACES.get(key)
    if ws is None or ws.numel() * 4 < nbytes:
        ws = torch.empty((max(nbytes, 256) + 3) // 4, dtype=torch.float32, device=device)
        _WORKSPACES[key] = ws
    return ws


# (saved_bytes, fwd_ws_bytes, bwd_ws_bytes, layout) per static configuration
_PLANS = {}


def _plan(call: WcLayerCall, key):
    pl = _PLANS.get(key)
    if pl is None:
        lay = WcSavedLayout()
        check(lib.wc_layer_saved_layout(ctypes.byref(call), ctypes.byref(lay)), 'wc_layer_saved_layout')
        pl = (int(lay.total), int(lib.wc_layer_workspace_bytes(ctypes.byref(call), 0)),
              int(lib.wc_layer_workspace_bytes(ctypes.byref(call), 1)), lay)
        _PLANS[key] = pl
    return pl


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
        Cout = next((w.shape[-1] for w in (Wu, Wc, D) if w is not None), Cin)
        E = class_matrix.shape[1] if class_matrix is not None else 0
        g = spec.group_size or Cin
        if spec.whiten and Cin % g:
            raise ValueError(f'channels {Cin} not divisible by group_size {g}')
        if spec.whiten and g > MAX_GROUP:
            raise ValueError(f'whitening group of {g} channels: the factorisation kernels hold one group matrix per '
                             f'CTA / cluster and support at most {MAX_GROUP}; pass group_size<={MAX_GROUP} '
                             f'(block-diagonal whitening) for C={Cin}')
        G = Cin // g
        # The kernels take raw device pointers: every operand is checked here (device, dtype, layout, shape), so a
        # mismatch raises instead of reading out of bounds.
        Wu, bu, Wc, bc = (_operand(t, n, dev, s) for t, n, s in (
            (Wu, 'Wu', (Cin, Cout)), (bu, 'bu', (Cout,)), (Wc, 'Wc', (K, Cin, Cout)), (bc, 'bc', (K, Cout))))
        class_matrix, D, bf = (_operand(t, n, dev, s) for t, n, s in (
            (class_matrix, 'class_matrix', (K, E)), (D, 'D', (E, Cin, Cout)), (bf, 'bf', (K, Cout))))
        if (class_matrix is None) != (D is None):
            raise ValueError('class_matrix and D (the FactorizedConv11 dictionary) go together')
        if spec.whiten:
            if moving_mean is None or moving_cov is None:
                raise ValueError('whitening needs moving_mean and moving_cov buffers')
            moving_mean = _operand(moving_mean, 'moving_mean', dev, None, numel=Cin, inplace=True)
            moving_cov = _operand(moving_cov, 'moving_cov', dev, None, numel=Cin * g, inplace=True)
        if cls is not None and cls.device != dev:
            raise ValueError(f'cls is on {cls.device}, x on {dev}: labels must live on the same CUDA device')
        if info is not None and (info.device != dev or info.dtype != torch.int32 or info.numel() < _lib.INFO_SLOTS):
            raise ValueError(f'info must be an int32 tensor of {_lib.INFO_SLOTS} slots on {dev}')
        dec = _lib.WC_DECOMP[spec.decomposition]
        world = sync.world_size(group)
        do_sync = spec.whiten and training and world > 1

        c = WcLayerCall()
        c.N, c.HW, c.Cin, c.Cout, c.g, c.K, c.E = N, HW, Cin, Cout, g, K, E
        c.whiten, c.decomposition, c.renorm = int(spec.whiten), dec, int(spec.renorm)
        c.training, c.triangular = int(training), int(spec.triangular)
        c.relu = int(spec.relu)
        c.use_tensor_cores = int(USE_TENSOR_CORES)
        c.eps, c.momentum = spec.epsilon, spec.momentum
        c.count = float(N * HW)
        c.Wu, c.bu, c.Wc, c.bc = ptr(Wu), ptr(bu), ptr(Wc), ptr(bc)
        c.class_matrix, c.D, c.bf = ptr(class_matrix), ptr(D), ptr(bf)
        c.moving_mean, c.moving_cov = ptr(moving_mean), ptr(moving_cov)
        c.cls, c.info = ptr(cls), ptr(info)
        has = tuple(t is not None for t in (Wu, bu, Wc, bc, class_matrix, D, bf))
        key = (N, HW, Cin, Cout, g, K, E, c.whiten, dec, c.renorm, c.training, c.triangular, c.use_tensor_cores, has)   # relu changes no size
        saved_bytes, fws, bws, lay = _plan(c, key)

        saved = torch.empty(saved_bytes // 4, dtype=torch.float32, device=dev)
        ws = _workspace(fws, dev, st, 0)
        y = _f32(N, H, W_, Cout, device=dev)
        c.x, c.y = ptr(x), ptr(y)
        c.saved, c.saved_bytes = ptr(saved), saved_bytes
        c.workspace, c.workspace_bytes = ptr(ws), ws.numel() * 4
        keep = [saved]
        xch = sync.peer_exchange(group, Cin, g, dev) if (do_sync and not spec.renorm) else None
        if xch is not None:
            # statistics exchanged over peer memory inside the call: one foreign call, no host collective
            c.xchg_peers, c.xchg_state = xch.peers_dev, ptr(xch.state)
            c.xchg_rank, c.xchg_world = xch.rank, xch.world
            c.phases = PHASE_STATS | PHASE_APPLY
            _call('wc_layer_fwd', ctypes.byref(c), st)
            keep.append(xch)
        elif do_sync:
            # phase 1: this shard's record {mean, S, count}; all-gather; Chan merge; phase 2 with the global statistics
            c.phases = PHASE_STATS
            _call('wc_layer_fwd', ctypes.byref(c), st)
            nrec = Cin + G * g * g + 1
            rec = saved[lay.rec // 4: lay.rec // 4 + nrec]
            rec[-1:].fill_(c.count)
            recs = sync.gather_shard_records(rec, group)
            merged = _f32(nrec, device=dev)     # mean | S | count of the concatenated batch
            _call('wc_stats_merge', ptr(recs), world, Cin, g, ptr(merged), ptr(merged[Cin:]), ptr(merged[nrec - 1:]), st)
            c.mean_ext, c.S_ext, c.count_dev = ptr(merged), ptr(merged[Cin:]), ptr(merged[nrec - 1:])
            keep.append(merged)
            c.phases = PHASE_APPLY
            _call('wc_layer_fwd', ctypes.byref(c), st)
        else:
            c.phases = PHASE_STATS | PHASE_APPLY
            _call('wc_layer_fwd', ctypes.byref(c), st)

        ctx.call, ctx.keep, ctx.group, ctx.sync = c, keep, group, (do_sync and xch is None)
        ctx.dims = (N, H, W_, Cin, Cout, g, K, E, bws)
        ctx.has = has
        ctx.stats_train = bool(spec.whiten and training)
        ctx.relu = bool(spec.relu)
        ctx.save_for_backward(x, cls, Wu, bu, Wc, bc, class_matrix, D, bf, moving_mean, y if spec.relu else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x = ctx.saved_tensors[0]
        c: WcLayerCall = ctx.call
        N, H, W_, Cin, Cout, g, K, E, bws = ctx.dims
        has_Wu, has_bu, has_Wc, has_bc, has_cm, has_D, has_bf = ctx.has
        need = ctx.needs_input_grad
        dev = x.device
        st = _stream()
        _need_cuda(dy, 'grad_output')
        if ctx.relu:
            # backward of the fused activation: one elementwise pass (dy, y -> dy); the kernels then see the gradient
            # with respect to the pre-activation output
            dy = dy * (ctx.saved_tensors[10] > 0)
        dy = dy.contiguous()
        G = Cin // g

        dWu = _f32(Cin, Cout, device=dev) if (has_Wu and need[1]) else None
        dbu = _f32(Cout, device=dev) if (has_bu and need[2]) else None
        dWc = _f32(K, Cin, Cout, device=dev) if (has_Wc and need[3]) else None
        dbc = _f32(K, Cout, device=dev) if (has_bc and need[4]) else None
        dcm = _f32(K, E, device=dev) if (has_cm and need[5]) else None
        dD = _f32(E, Cin, Cout, device=dev) if (has_D and need[6]) else None
        dbf = _f32(K, Cout, device=dev) if (has_bf and need[7]) else None
        dx = torch.empty_like(x) if need[0] else None
        stats_grad = ctx.stats_train and need[0]
        red = _f32(G * g * g + Cin, device=dev) if stats_grad else None
        ws = _workspace(bws, dev, st, 1)

        c.dy, c.dx = ptr(dy), ptr(dx)
        c.dWu, c.dbu, c.dWc, c.dbc = ptr(dWu), ptr(dbu), ptr(dWc), ptr(dbc)
        c.dclass_matrix, c.dD, c.dbf = ptr(dcm), ptr(dD), ptr(dbf)
        c.red = ptr(red)
        c.workspace, c.workspace_bytes = ptr(ws), ws.numel() * 4
        if ctx.sync and stats_grad:
            c.phases = PHASE_REDUCE
            _call('wc_layer_bwd', ctypes.byref(c), st)
            sync.allreduce_adjoint(red, ctx.group)
            c.phases = PHASE_DX
            _call('wc_layer_bwd', ctypes.byref(c), st)
        else:
            c.phases = PHASE_REDUCE | PHASE_DX
            _call('wc_layer_bwd', ctypes.byref(c), st)
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
