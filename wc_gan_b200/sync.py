"""Synchronised-batch whitening: the two small exchanges per WC layer.

forward : every rank contributes one shard record  {mean[C], S[G*g*g], count}  (centred scatter of ITS rows);
          one all-gather, then libwc_b200's wc_stats_merge applies Chan's parallel update so that mean / covariance
          equal those of the concatenated batch (gan/conditional_layers.py:658,:683 on the global batch).
backward: the adjoint of the whitening matrix and of the mean, {dWeff[G*g*g], gbar[C]}, is all-reduced (sum) before
          the factorisation adjoint, so A and cvec are global; coloring-parameter gradients stay local, as with
          torch.nn.SyncBatchNorm (DDP averages them).

Only the collectives live here (device-agnostic torch.distributed calls, NCCL on GPUs, gloo in the CPU tests);
the arithmetic on either side of them runs in the CUDA library.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def record_numel(C: int, g: int) -> int:
    return C + (C // g) * g * g + 1


def split_record(rec: torch.Tensor, C: int, g: int):
    """Views (mean, S, count) into a packed shard record."""
    n = (C // g) * g * g
    return rec[..., :C], rec[..., C:C + n], rec[..., C + n:]


def world_size(group) -> int:
    if group is None or not dist.is_available() or not dist.is_initialized():
        return 1
    return dist.get_world_size(group)


def gather_shard_records(rec: torch.Tensor, group) -> torch.Tensor:
    """(L,) -> (R, L): every rank's record, in rank order."""
    R = dist.get_world_size(group)
    out = torch.empty(R, rec.numel(), dtype=rec.dtype, device=rec.device)
    try:
        dist.all_gather_into_tensor(out, rec.contiguous(), group=group)
    except (RuntimeError, NotImplementedError):   # backends without the flat variant
        parts = [torch.empty_like(rec) for _ in range(R)]
        dist.all_gather(parts, rec.contiguous(), group=group)
        out = torch.stack(parts)
    return out


def allreduce_adjoint(red: torch.Tensor, group) -> torch.Tensor:
    """In-place sum over ranks of the packed {dWeff, gbar} buffer."""
    dist.all_reduce(red, op=dist.ReduceOp.SUM, group=group)
    return red


def merge_records_reference(recs: torch.Tensor, C: int, g: int):
    """Chan merge in plain torch (what wc_stats_merge computes on the device); used by the CPU tests."""
    mean_r, S_r, n_r = split_record(recs.double(), C, g)
    n = n_r.sum()
    mean = (n_r * mean_r).sum(0) / n
    G = C // g
    d = (mean_r - mean).reshape(-1, G, g)
    S = (S_r.reshape(-1, G, g, g) + n_r.reshape(-1, 1, 1, 1) * d[:, :, :, None] * d[:, :, None, :]).sum(0)
    return mean, S, n
