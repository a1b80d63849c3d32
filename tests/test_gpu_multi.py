"""Synchronised-batch whitening on >= 2 GPUs (NCCL): an R-way sharded run equals the 1-GPU run on the
concatenated batch.  Skipped when fewer than 2 devices are visible."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q, shape, after_norm, K, E, transport):
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    os.environ['WC_B200_SYNC'] = transport
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        import wc_gan_b200 as W
        from oracle import wc_oracle as O
        dev = torch.device('cuda', rank)
        x = O.make_activation(shape, 'correlated', seed=0)
        gen = torch.Generator().manual_seed(1)
        cls = torch.randint(0, K, (shape[0], 1), generator=gen, dtype=torch.int32)
        dy = torch.randn(shape, generator=gen)
        per = shape[0] // world
        sl = slice(rank * per, (rank + 1) * per)
        torch.manual_seed(0)
        stack = W.WhiteningColoring('d', after_norm, number_of_classes=K, filters_emb=E, name='s', process_group=dist.group.WORLD)
        stack.build(shape[-1], dev); stack.built = True
        stack.to(dev).train()
        xs = x[sl].to(dev).requires_grad_(True)
        for _ in range(3):        # three exchanges per direction: both data slots and the epoch counters are reused
            xs.grad = None
            y = stack(xs, cls[sl].to(dev))
            y.backward(dy[sl].to(dev))
        torch.cuda.synchronize()
        from wc_gan_b200 import sync
        used_peer = any(v is not None for v in sync._EXCHANGES.values())
        q.put((rank, y.detach().cpu(), xs.grad.cpu(), stack.npart.moving_variance.cpu(), used_peer))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('transport', ['peer', 'nccl'])
@pytest.mark.parametrize('shape,after_norm,K,E', [((8, 8, 8, 128), 'ufconv', 5, 3), ((4, 16, 16, 256), 'uconv', 5, 3),
                                                   ((6, 4, 4, 24), 'ucconv', 5, 3), ((16, 4, 4, 128), 'ufconv', 40, 4)])
def test_sharded_equals_single(shape, after_norm, K, E, transport):
    """transport 'peer': statistics and adjoint exchanged inside the library's kernels over symmetric peer memory
    (wc_exchange.cu); 'nccl': host-issued all-gather / all-reduce between the phases."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    import wc_gan_b200 as W
    from oracle import wc_oracle as O
    # (K = 40, E = 4: the factorised coloring forms; with the peer transport the adjoint all-reduce is a stage of that
    # chain launch, with NCCL the split-phase backward rebuilds the W_k the forward did not form)
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, shape, after_norm, K, E, transport)) for r in range(2)]
    for p in procs:
        p.start()
    out = {}
    for _ in range(2):
        r = q.get(timeout=300)
        out[r[0]] = r[1:]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-GPU run on the whole batch with identically seeded parameters
    dev = torch.device('cuda', 0)
    x = O.make_activation(shape, 'correlated', seed=0)
    gen = torch.Generator().manual_seed(1)
    cls = torch.randint(0, K, (shape[0], 1), generator=gen, dtype=torch.int32)
    dy = torch.randn(shape, generator=gen)
    torch.manual_seed(0)
    stack = W.WhiteningColoring('d', after_norm, number_of_classes=K, filters_emb=E, name='s')
    stack.build(shape[-1], dev); stack.built = True
    stack.to(dev).train()
    xs = x.to(dev).requires_grad_(True)
    for _ in range(3):
        xs.grad = None
        y = stack(xs, cls.to(dev)); y.backward(dy.to(dev))
    if transport == 'nccl':
        assert not out[0][3] and not out[1][3]
    elif not (out[0][3] and out[1][3]):
        pytest.skip('symmetric peer memory is not available on this box: the peer transport fell back to NCCL')
    ysh = torch.cat([out[0][0], out[1][0]]); dxsh = torch.cat([out[0][1], out[1][1]])
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
    # both runs are within the 1e-4 / 1e-3 parity tolerance of the fp64 oracle; they differ from each other only by
    # fp32 summation order (different shard pivots, Chan merge), which the whitening conditioning amplifies
    assert rel(ysh, y.detach().cpu()) < 1e-4
    assert rel(dxsh, xs.grad.cpu()) < 1e-3
    assert rel(out[0][2], stack.npart.moving_variance.cpu()) < 1e-5 and torch.equal(out[0][2], out[1][2])
