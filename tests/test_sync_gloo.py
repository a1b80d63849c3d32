"""Synchronised-batch whitening protocol over real torch.distributed collectives (gloo, world_size 2, CPU).

Each rank holds a shard, builds its shard record {mean, S, count} with the oracle, exchanges it with
wc_gan_b200.sync (the same functions the CUDA path calls over NCCL), merges, runs the closed-form forward /
backward with the merged statistics and the all-reduced adjoint, and the concatenated result must equal the
single-process full-batch oracle (SURVEY.md section 8e "Check")."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import closed_form as CF
from oracle import wc_oracle as O

C, G_SIZE, K, E = 8, 4, 3, 2
SHAPE = (6, 3, 2, C)
SPLIT = 2   # rank 0 gets samples [:2], rank 1 gets [2:]  (unequal shards on purpose)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _inputs():
    x = O.make_activation(SHAPE, 'relu', seed=7, dtype=torch.float64)
    gen = torch.Generator().manual_seed(3)
    cls = torch.randint(0, K, (SHAPE[0], 1), generator=gen, dtype=torch.int32)
    dy = torch.randn(SHAPE, generator=gen, dtype=torch.float64)
    params = O.make_params(C, after_norm='ufconv', number_of_classes=K, filters_emb=E, group_size=G_SIZE, seed=1,
                           dtype=torch.float64, random_bias=True)
    return x, cls, dy, params


def _worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from wc_gan_b200 import sync
        x, cls, dy, params = _inputs()
        sl = slice(0, SPLIT) if rank == 0 else slice(SPLIT, None)
        xs, cs, dys = x[sl], cls[sl], dy[sl]
        X = xs.reshape(-1, C)
        mu = X.mean(0)
        Xc = X - mu
        G = C // G_SIZE
        S = torch.stack([Xc[:, i * G_SIZE:(i + 1) * G_SIZE].t() @ Xc[:, i * G_SIZE:(i + 1) * G_SIZE] for i in range(G)])
        rec = torch.cat([mu, S.reshape(-1), torch.tensor([float(X.shape[0])], dtype=torch.float64)])
        assert rec.numel() == sync.record_numel(C, G_SIZE)
        recs = sync.gather_shard_records(rec, dist.group.WORLD)
        assert recs.shape == (world, rec.numel())
        mean, Sg, n = sync.merge_records_reference(recs, C, G_SIZE)
        kw = dict(norm='d', after_norm='ufconv', group_size=G_SIZE, number_of_classes=K)
        fwd = CF.forward(xs, cs, params, stats_override=(float(n), mean, Sg), **kw)
        # local reductions -> all-reduce of {dWeff, gbar} -> global adjoint
        b0 = CF.backward(xs, dys, fwd, params, norm='d', after_norm='ufconv')
        red = torch.cat([b0['_inter']['dWeff'].reshape(-1), b0['_inter']['gbar']])
        sync.allreduce_adjoint(red, dist.group.WORLD)
        ng = G * G_SIZE * G_SIZE
        bwd = CF.backward(xs, dys, fwd, params, norm='d', after_norm='ufconv',
                          reduced_override=(red[:ng].reshape(G, G_SIZE, G_SIZE), red[ng:]))
        gk = bwd['u_kernel'].clone()
        dist.all_reduce(gk)     # parameter gradients are summed by the data-parallel wrapper (DDP), not by the layer
        # numpy copies: torch tensors travel through the queue as shared-memory files that vanish when the worker exits
        q.put((rank,) + tuple(t.detach().numpy().copy() for t in
                              (fwd['y'], bwd['x'], gk, fwd['new_moving_mean'], fwd['new_moving_variance'])))
    finally:
        dist.destroy_process_group()


def test_two_rank_sync_batch_matches_full_batch():
    x, cls, dy, params = _inputs()
    kw = dict(norm='d', after_norm='ufconv', group_size=G_SIZE, number_of_classes=K)
    full = CF.forward(x, cls, params, **kw)
    fb = CF.backward(x, dy, full, params, norm='d', after_norm='ufconv')
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = {}
    for _ in range(2):
        r = q.get(timeout=120)
        out[r[0]] = tuple(torch.from_numpy(a) for a in r[1:])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    y = torch.cat([out[0][0], out[1][0]])
    dx = torch.cat([out[0][1], out[1][1]])
    assert torch.allclose(y, full['y'], rtol=1e-10, atol=1e-10)
    assert torch.allclose(dx, fb['x'], rtol=1e-9, atol=1e-10)
    assert torch.allclose(out[0][2], fb['u_kernel'], rtol=1e-9, atol=1e-10)
    for r in (0, 1):   # moving statistics identical on every rank and equal to the full-batch update
        assert torch.allclose(out[r][3], full['new_moving_mean'], rtol=1e-12, atol=1e-12)
        assert torch.allclose(out[r][4], full['new_moving_variance'], rtol=1e-12, atol=1e-12)


def test_record_layout_matches_library_contract():
    from wc_gan_b200 import sync
    rec = torch.arange(sync.record_numel(8, 4), dtype=torch.float32)
    m, S, n = sync.split_record(rec, 8, 4)
    assert m.numel() == 8 and S.numel() == 2 * 16 and n.numel() == 1 and float(n) == rec.numel() - 1
