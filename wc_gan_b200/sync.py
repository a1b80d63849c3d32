"""Synchronised-batch whitening: the two small exchanges per WC layer.

forward : every rank contributes one shard record  {mean[C], S[G*g*g], count}  (centred scatter of ITS rows);
          one all-gather, then libwc_b200's wc_stats_merge applies Chan's parallel update so that mean / covariance
          equal those of the concatenated batch (gan/conditional_layers.py:658,:683 on the global batch).
backward: the adjoint of the whitening matrix and of the mean, {dWeff[G*g*g], gbar[C]}, is all-reduced (sum) before
          the factorisation adjoint, so A and cvec are global; coloring-parameter gradients stay local, as with
          torch.nn.SyncBatchNorm (DDP averages them).

Two transports.  On one node with peer access the exchanges run INSIDE the library's kernels over a symmetric
buffer (`peer_exchange`: torch's symmetric memory only allocates and maps the buffer; wc_exchange.cu pushes, signals,
waits and merges -- no NCCL call, no host round trip, one foreign call per direction).  Otherwise the host issues the
collectives between the two phases of each direction (device-agnostic torch.distributed calls: NCCL on GPUs, gloo in
the CPU tests); the arithmetic on either side of them runs in the CUDA library.
"""
from __future__ import annotations

import os
import warnings

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


class PeerExchange:
    """A symmetric exchange buffer mapped on every rank of `group`, plus this rank's epoch counters."""

    def __init__(self, buf, handle, state, rank, world):
        self.buf, self.handle, self.state, self.rank, self.world = buf, handle, state, rank, world
        # device array of `world` buffer pointers, rank order.  A copy in ordinary device memory: the kernels read it on
        # their serial path, and the allocator's own table (handle.buffer_ptrs_dev) measured ~17 us per exchange to read
        # (scratch/xchtime.py: "pushed" 23 us after the kernel start with or without remote stores)
        ptrs = [int(p) for p in handle.buffer_ptrs]
        self.peer_table = torch.tensor(ptrs, dtype=torch.int64, device=buf.device)
        self.peers_dev = self.peer_table.data_ptr()


_EXCHANGES = {}      # (process group object, C, g, device index) -> PeerExchange | None
_WARNED = False


def _all_agree(ok: bool, pg, device) -> bool:
    """True iff `ok` on every rank (one small all-reduce: every rank calls it at the same point, whatever happened
    to it locally, so nobody is left waiting in a collective the others skipped)."""
    flag = torch.tensor([1 if ok else 0], device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=pg)
    return int(flag.item()) == 1


def peer_exchange(group, C: int, g: int, device):
    """The peer-memory transport for layers of shape (C, g) on `group`, or None (then the host runs NCCL collectives).
    Collective on first use per (group, C, g): every rank reaches it at the same layer.

    Requirements of this transport (the NCCL path has none of them; WC_B200_SYNC=nccl forces it):
      * layers of the same (C, g) share one buffer and one epoch counter, so their exchanges must be enqueued on ONE
        stream, in the same order on every rank (true for a model stepped the same way on every rank);
      * the exchange kernels spin until every peer has pushed: a kernel that a peer enqueues before ITS exchange must
        not wait on this rank (e.g. an NCCL collective issued from an autograd hook between the layer's kernels on one
        rank only).  With DDP keep the default CUDA_DEVICE_MAX_CONNECTIONS so NCCL's kernels and these run concurrently;
      * a peer that never arrives (crashed rank) is detected after WC_B200_XCHG_TIMEOUT_MS (default 10000): the kernel
        gives up, sets info[WC_INFO_XCHG_TIMEOUT] and the layer's check_info() raises; results of that step are garbage.
    """
    global _WARNED
    if os.environ.get('WC_B200_SYNC', '').lower() == 'nccl' or device.type != 'cuda':
        return None
    pg = group if group is not None else dist.group.WORLD
    key = (pg, C, g, device.index)          # the group object itself: id() can be reused after garbage collection
    if key in _EXCHANGES:
        return _EXCHANGES[key]
    # Every fallible step is followed by an agreement round, so a rank that failed locally never skips a collective
    # the others are already blocked in (symm.rendezvous and dist.barrier are collectives).
    xch, err = None, None
    buf = handle = None
    try:
        import torch.distributed._symmetric_memory as symm
        from ._lib import lib
        world, rank = dist.get_world_size(pg), dist.get_rank(pg)
        nbytes = int(lib.wc_exchange_bytes(C, g, world))
        if nbytes == 0:
            raise RuntimeError(f'no peer exchange for C={C}, g={g}, world={world}')
        buf = symm.empty((nbytes + 3) // 4, dtype=torch.float32, device=device)
        buf.zero_()
    except Exception as e:                  # no symmetric memory here
        err = e
    if _all_agree(err is None, pg, device):
        try:
            handle = symm.rendezvous(buf, pg)
            state = torch.zeros(4, dtype=torch.int32, device=device)
            torch.cuda.synchronize(device)
        except Exception as e:              # no P2P mapping between these devices
            err = e
        if _all_agree(err is None, pg, device):
            dist.barrier(group=pg)          # every buffer is zeroed before anyone's first push can land in it
            xch = PeerExchange(buf, handle, state, rank, world)
    if xch is None and not _WARNED:
        warnings.warn(f'wc_gan_b200: peer-memory exchange unavailable ({err!r} on this or another rank); '
                      f'using NCCL collectives')
        _WARNED = True
    _EXCHANGES[key] = xch
    return xch
