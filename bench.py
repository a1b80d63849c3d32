#!/usr/bin/env python
"""bench.py -- WC layer fwd+bwd throughput on B200 (BASELINE.json's metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N ...            # the reference's CPU path (oracle port)

A "step" is one forward + backward of the whitening-coloring layer on one synthetic
(256,16,16,256) fp32 NHWC activation per GPU (BASELINE.json north_star's headline shape; the
reference's default stack ``create_norm('d','uconv')``: Cholesky whitening over all C channels +
unconditional coloring).  value = algorithmic bytes (32*M*C per step per GPU)
of all ranks / max-over-ranks device time, in GB/s.  With N>1 every rank holds its own shard of the
same size (weak scaling) and the layer runs synchronised-batch whitening over NCCL.

One JSON line is printed by rank 0.  See DESIGN.md "Measurement" for every key.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SHAPE = (256, 16, 16, 256)
METRIC = 'WC layer fwd+bwd GB/s'
UNIT = 'GB/s'


def alg_bytes(shape) -> int:
    n, h, w, c = shape
    return 32 * n * h * w * c          # 8 tensor passes x 4 B  


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return float(d['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def workload_config(n_gpus: int) -> dict:
    return {'workload': 'cfg3: WC layer fwd+bwd on 256x16x16x256 fp32 NHWC per GPU',
            'stack': "create_norm('d','uconv'): Cholesky whitening g=C=256 + unconditional coloring",
            'distribution': 'N(1, 0.5 I + 0.5 11^T) (gan/unit_tests.py:621-623)', 'eps': 1e-3, 'momentum': 0.99,
            'phase': 'training', 'parallelism': f'dp{n_gpus} sync-batch whitening' if n_gpus > 1 else 'single GPU',
            'l2': 'inputs rotate over 4 buffer sets (x, dy: 67 MB each per set; 268 MB touched per step) > 2x L2'}


# --------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md recipe)
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '100', '-i', str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float) -> dict:
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        rows = [r for r in self.rows if t0 - 0.05 <= r[0] <= t1 + 0.15] or self.rows
        for _, line in rows:
            f = [x.strip() for x in line.split(',')]
            try:
                sm.append(float(f[1])); smax = float(f[2])
            except Exception:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': smax,
                'reasons': sorted(reasons), 'samples': len(sm)}


# --------------------------------------------------------------------------------------------------
# CPU reference (the oracle port of the reference's TF graph; TF itself is not installable here)
# --------------------------------------------------------------------------------------------------
def cpu_reference_time(shape, steps: int, warmup: int):
    import torch
    from oracle import wc_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    x = O.make_activation(shape, 'correlated', seed=0)
    params = O.make_params(shape[-1], after_norm='uconv', seed=3)
    leaves = {k: v.clone().requires_grad_(True) for k, v in params.items() if k.startswith('u_')}
    p = dict(params); p.update(leaves)
    dy = torch.randn(shape)
    ts = []
    for i in range(warmup + steps):
        xr = x.clone().requires_grad_(True)
        t0 = time.perf_counter()
        y, mm, mc = O.wc_block(xr, None, p, norm='d', after_norm='uconv', training=True)
        torch.autograd.grad(y, [xr] + list(leaves.values()), dy)
        t1 = time.perf_counter()
        if i >= warmup:
            ts.append(t1 - t0)
    return ts, cores


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 20))
    ts, cores = cpu_reference_time(SHAPE, steps, max(1, min(args.warmup, 3)))
    total = sum(ts)
    val = alg_bytes(SHAPE) * len(ts) / total / 1e9
    sample = f'{len(ts)} full fwd+bwd steps of the same 256x16x16x256 workload (one shard), torch-CPU fp32 port of the TF graph, {cores} threads'
    out = {'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': len(ts),
           'warmup': max(1, min(args.warmup, 3)), 'ms_per_step': 1e3 * total / len(ts), 'higher_is_better': True,
           'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
           'config': workload_config(args.gpus),
           'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
           'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
           'note': 'TensorFlow 1.5 / Keras 2.0.8 cannot be installed here (no network, Python 3.12): the reference arm is '
                   'oracle/wc_oracle.py, the op-for-op CPU restatement of gan/conditional_layers.py:646-712 + generator.py:49-51'}
    print(json.dumps(out), flush=True)


# --------------------------------------------------------------------------------------------------
# this repo's arm
# --------------------------------------------------------------------------------------------------
def make_inputs(torch, shape, nbuf, device, seed):
    g = torch.Generator(device=device).manual_seed(seed)
    xs, dys = [], []
    n, h, w, c = shape
    for _ in range(nbuf):
        z = torch.randn(shape, generator=g, device=device)
        zc = torch.randn((n, h, w, 1), generator=g, device=device)
        xs.append((z * 0.5 ** 0.5 + zc * 0.5 ** 0.5 + 1.0).contiguous())
        dys.append(torch.randn(shape, generator=g, device=device))
    return xs, dys


def build_stack(W, torch, norm, after_norm, c, device, *, dec='cholesky', g=None, K=None, E=10, group=None):
    s = W.WhiteningColoring(norm, after_norm, number_of_classes=K, filters_emb=E, name='bench', decomposition=dec,
                            group_size=g, process_group=group)
    s.build(c, device); s.built = True
    s.to(device).train()
    return s


def time_steps(torch, fn, steps, warmup, barrier=None):
    for i in range(warmup):
        fn(i)
    if barrier:
        barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    for i in range(steps):
        fn(warmup + i)
    e1.record()
    if barrier:
        barrier()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1), t0, time.time()


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: this repository has no CPU path (use --impl reference for the CPU baseline)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    group = None
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
        group = dist.group.WORLD
    import wc_gan_b200 as W
    from wc_gan_b200._lib import lib

    barrier = (lambda: dist.barrier()) if world > 1 else None
    NBUF = 4
    xs, dys = make_inputs(torch, SHAPE, NBUF, dev, seed=1234 + rank)
    stack = build_stack(W, torch, 'd', 'uconv', SHAPE[-1], dev, group=group)
    params = [p for p in stack.parameters()]

    def eager_step(i):
        x = xs[i % NBUF].requires_grad_(True)
        y = stack(x)
        return y, torch.autograd.grad(y, [x] + params, dys[i % NBUF])

    # The step is launch-bound on the host (about 25 small launches per layer), so by default it is captured once per
    # input buffer set in a CUDA graph and replayed; --no-graph times the eager path.  Same kernels, same work.
    graphs = []
    if not args.no_graph:
        for i in range(3):
            eager_step(i)
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for i in range(NBUF):
                eager_step(i)                      # warm the capture stream's workspaces
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        for i in range(NBUF):
            gph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gph, stream=side):
                out = eager_step(i)
            graphs.append((gph, out))

    def step(i):
        if graphs:
            graphs[i % NBUF][0].replay()
            return graphs[i % NBUF][1]
        return eager_step(i)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = lib.wc_launch_count()
    for i in range(args.warmup):
        step(i)
    l1 = lib.wc_launch_count()
    ms, t0, t1 = time_steps(torch, step, args.steps, 0, barrier)
    launches = lib.wc_launch_count() - l1
    if graphs:   # replays do not pass through the library's counter: count one eager step
        lc = lib.wc_launch_count(); eager_step(0); launches = (lib.wc_launch_count() - lc) * args.steps
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    per_step = ms / args.steps
    value = world * alg_bytes(SHAPE) / (per_step * 1e-3) / 1e9
    hbm, peak_src = peaks()

    # ---- per-stage device time: the four M x C stages launched standalone through the fine-grained C ABI on the
    #      same rotating buffers, CUDA events on the launching stream around `n` back-to-back launches
    stages = stage_times(torch, lib, xs, dys, min(max(args.steps, 10), 50))
    M, C = SHAPE[0] * SHAPE[1] * SHAPE[2], SHAPE[3]
    stage_bytes = {'wc_stats_fwd': 4 * M * C, 'wc_apply_fwd': 8 * M * C, 'wc_reduce_bwd': 8 * M * C, 'wc_apply_bwd': 12 * M * C}
    dom = max(stage_bytes, key=lambda k: stages.get(k, 0.0))
    ach = stage_bytes[dom] / (stages[dom] * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(dom)
        except Exception:
            traffic = None
    roofline = {'bound': 'hbm', 'kernel': dom, 'achieved': ach, 'peak': hbm, 'unit': 'GB/s', 'frac': ach / hbm,
                'traffic': traffic, 'peak_source': peak_src, 'algorithmic_bytes_per_launch': stage_bytes[dom],
                'avg_launch_ms': stages[dom]}

    out = None
    if rank == 0:
        out = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
               'ms_per_step': per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
               'data': 'synthetic', 'config': workload_config(world), 'roofline': roofline,
               'hbm_frac_fwd_bwd': value / world / hbm, 'hbm_frac_fwd_bwd_single_touch': value / world / hbm / 2,
               'stages_ms': stages, 'gpu_launches': int(launches), 'clocks': clocks,
               'path': int(lib.wc_path_for(M, C, C, C)), 'cuda_graph': bool(graphs)}

    # ---- end to end through the public module with HOST buffers (H2D + D2H inside the timed region)
    # Every step copies ITS inputs from pinned host memory and ITS results back; the three legs run on three streams
    # with two buffer sets, so step i's read-back overlaps step i+1's upload (PCIe is full duplex) and both overlap
    # the kernels of the step in between.
    xh = [x.cpu().pin_memory() for x in xs[:2]]
    dyh = [d.cpu().pin_memory() for d in dys[:2]]
    yh = [torch.empty(SHAPE, dtype=torch.float32).pin_memory() for _ in range(2)]
    dxh = [torch.empty(SHAPE, dtype=torch.float32).pin_memory() for _ in range(2)]
    xd = [torch.empty(SHAPE, device=dev) for _ in range(2)]
    dyd = [torch.empty(SHAPE, device=dev) for _ in range(2)]
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_comp = [torch.cuda.Event() for _ in range(2)]
    ev_out = [torch.cuda.Event() for _ in range(2)]
    alive = [None, None]
    cur = torch.cuda.current_stream()
    for e in ev_comp + ev_out:
        e.record(cur)

    def e2e_step(i):
        b = i % 2
        with torch.cuda.stream(s_in):
            s_in.wait_event(ev_comp[b])            # the step that last read this input pair has run
            xd[b].copy_(xh[b], non_blocking=True)
            dyd[b].copy_(dyh[b], non_blocking=True)
            ev_in[b].record(s_in)
        cur.wait_event(ev_in[b])
        xi = xd[b].detach().requires_grad_(True)
        y = stack(xi)
        gr = torch.autograd.grad(y, [xi] + params, dyd[b])
        ev_comp[b].record(cur)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_comp[b])
            s_out.wait_event(ev_out[b])            # the previous read-back into this host pair is done
            yh[b].copy_(y.detach(), non_blocking=True)
            dxh[b].copy_(gr[0], non_blocking=True)
            ev_out[b].record(s_out)
        y.record_stream(s_out); gr[0].record_stream(s_out)
        alive[b] = (y, gr)

    def e2e_drain():
        cur.wait_stream(s_in); cur.wait_stream(s_out)

    e_steps = max(4, min(args.steps, 10))
    for i in range(2):
        e2e_step(i)
    e2e_drain()
    if barrier:
        barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(cur)
    for i in range(e_steps):
        e2e_step(2 + i)
    e2e_drain()
    e1.record(cur)
    if barrier:
        barrier()
    torch.cuda.synchronize()
    ems = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ems], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ems = float(t.item())
    nbytes = 4 * M * C
    if rank == 0:
        out['e2e'] = {'value': world * alg_bytes(SHAPE) / (ems / e_steps * 1e-3) / 1e9, 'unit': UNIT,
                      'h2d_bytes_per_step': 2 * nbytes, 'd2h_bytes_per_step': 2 * nbytes, 'ms_per_step': ems / e_steps,
                      'steps': e_steps,
                      'what': 'every step: pinned host x, dy -> device; module fwd+bwd; y, dx -> pinned host; upload, '
                              'kernels and read-back of consecutive steps overlap on three streams (two buffer sets)'}

    # ---- other configurations (parity-test cases; reported, not the headline)
    if rank == 0 and world == 1 and not args.no_sweep:
        del xs[1:], dys[1:]
        out['sweep'] = sweep(torch, W, dev, hbm)

    # ---- CPU baseline beside it (rank 0, N == 1 only)
    if rank == 0 and world == 1 and not args.no_cpu:
        ts, cores = cpu_reference_time(SHAPE, 5, 1)
        v = alg_bytes(SHAPE) * len(ts) / sum(ts) / 1e9
        out['cpu_baseline'] = {'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'ms_per_step': 1e3 * sum(ts) / len(ts),
                               'sample': f'{len(ts)} full fwd+bwd steps of the same 256x16x16x256 workload; oracle/wc_oracle.py '
                                         f'(torch-CPU fp32 restatement of the TF graph), {cores} threads'}
    elif rank == 0:
        out['cpu_baseline'] = None
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        # The captured graphs hold NCCL work; tearing the communicator down under them can block for minutes.  Drop
        # the graphs, drain the device, meet the other ranks once more, then leave without the NCCL teardown.
        graphs.clear()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush(); sys.stderr.flush()
        os._exit(0)


def stage_times(torch, lib, xs, dys, n):
    """ms per launch of K1 (wc_stats_fwd), K3 (wc_apply_fwd), K4 (wc_reduce_bwd), K6 (wc_apply_bwd) at the bench shape.
    Each stage includes its helper launches (memset, finalize).  Operand values do not matter for the timing."""
    from wc_gan_b200._lib import ptr
    N, H, Wd, C = SHAPE
    HW, M = H * Wd, N * H * Wd
    dev = xs[0].device
    st = torch.cuda.current_stream().cuda_stream
    f = lambda *sh: torch.zeros(sh, dtype=torch.float32, device=dev)
    mean, S, T, b, A, cvec = f(C), f(C, C), torch.randn(C, C, device=dev) / 16, f(C), torch.randn(C, C, device=dev) / 16, f(C)
    nb = lib.wc_stats_workspace_bytes(M, C, C)
    ws = f(nb // 4 + 4)
    Ps = f(C * C + C)
    y = torch.empty_like(xs[0])
    Tp, TAp = f(lib.wc_pack_bytes(C, 1, 0) // 4), f(lib.wc_pack_bytes(C, 1, 1) // 4)
    assert lib.wc_pack_fwd(ptr(T), C, 1, ptr(Tp), st) == 0 and lib.wc_pack_bwd(ptr(T), ptr(A), C, C, 1, ptr(TAp), st) == 0
    nbuf = len(xs)
    fns = {
        'wc_stats_fwd': lambda i: lib.wc_stats_fwd(ptr(xs[i % nbuf]), M, C, C, ptr(ws), nb, ptr(mean), ptr(S), st),
        'wc_apply_fwd': lambda i: lib.wc_apply_fwd(ptr(xs[i % nbuf]), N, HW, C, C, ptr(mean), ptr(T), ptr(Tp), ptr(b), None, 1, 0,
                                                   ptr(y), None, st),
        'wc_reduce_bwd': lambda i: lib.wc_reduce_bwd(ptr(xs[i % nbuf]), ptr(dys[i % nbuf]), N, HW, C, C, ptr(mean), None, 1,
                                                     ptr(Ps), ptr(Ps[C * C:]), None, 0, st),
        'wc_apply_bwd': lambda i: lib.wc_apply_bwd(ptr(xs[i % nbuf]), ptr(dys[i % nbuf]), N, HW, C, C, ptr(mean), ptr(T), ptr(A),
                                                   ptr(TAp), ptr(cvec), C, None, 1, ptr(y), st),
    }
    out = {}
    for name, fn in fns.items():
        for i in range(3):
            assert fn(i) == 0, name
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        out[name] = e0.elapsed_time(e1) / n
    return out


def sweep(torch, W, dev, hbm):
    """cfg1 / cfg2 / cfg3 variants: ms per fwd+bwd, effective GB/s on 32*M*C, fraction of the HBM peak."""
    cases = [
        ('cfg1 64x8x8x256 chol g=256 uconv (L2-resident)', (64, 8, 8, 256), 'd', 'uconv', 'cholesky', None, None, 0),
        ('cfg2 64x16x16x128 chol ucconv K=10 (L2-resident)', (64, 16, 16, 128), 'd', 'ucconv', 'cholesky', None, 10, 0),
        ('cfg2 64x16x16x128 chol ufconv K=10 E=4 (L2-resident)', (64, 16, 16, 128), 'd', 'ufconv', 'cholesky', None, 10, 4),
        ('cfg3 256x16x16x256 chol g=16 uconv', SHAPE, 'd', 'uconv', 'cholesky', 16, None, 0),
        ('cfg3 256x16x16x256 chol g=32 uconv', SHAPE, 'd', 'uconv', 'cholesky', 32, None, 0),
        ('cfg3 256x16x16x256 chol g=64 uconv', SHAPE, 'd', 'uconv', 'cholesky', 64, None, 0),
        ('cfg3 256x16x16x256 zca g=16 uconv', SHAPE, 'd', 'uconv', 'zca', 16, None, 0),
        ('cfg3 256x16x16x256 zca g=32 uconv', SHAPE, 'd', 'uconv', 'zca', 32, None, 0),
        ('cfg3 256x16x16x256 zca g=64 uconv', SHAPE, 'd', 'uconv', 'zca', 64, None, 0),
        ('cfg3 256x16x16x256 zca g=256 uconv', SHAPE, 'd', 'uconv', 'zca', None, None, 0),
    ]
    res = []
    for name, shape, norm, an, dec, g, K, E in cases:
        nbuf = 4 if shape == SHAPE else 8
        xs, dys = make_inputs(torch, shape, nbuf if shape != SHAPE else 2, dev, seed=7)
        st = build_stack(W, torch, norm, an, shape[-1], dev, dec=dec, g=g, K=K, E=E or 10)
        params = [p for p in st.parameters()]
        cls = torch.randint(0, K, (shape[0], 1), dtype=torch.int32, device=dev) if K else None

        def step(i, xs=xs, dys=dys, st=st, params=params, cls=cls):
            x = xs[i % len(xs)].requires_grad_(True)
            y = st(x, cls)
            torch.autograd.grad(y, [x] + params, dys[i % len(xs)])
        steps = 5 if (dec == 'zca' and g is None) else 20
        # like the headline, the step is replayed from CUDA graphs (one per buffer set): the small cases are otherwise
        # bound by the host's ~15 us per launch, not by the device
        graphs = []
        try:
            for i in range(3):
                step(i)
            torch.cuda.synchronize()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for i in range(len(xs)):
                    step(i)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            for i in range(len(xs)):
                gph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gph, stream=side):
                    step(i)
                graphs.append(gph)
        except Exception:
            graphs = []
            torch.cuda.synchronize()
        run = (lambda i, graphs=graphs: graphs[i % len(graphs)].replay()) if graphs else step
        ms, _, _ = time_steps(torch, run, steps, 3)
        per = ms / steps
        gbs = alg_bytes(shape) / (per * 1e-3) / 1e9
        res.append({'case': name, 'ms_per_step': per, 'gbs': gbs, 'hbm_frac': gbs / hbm, 'cuda_graph': bool(graphs)})
        del xs, dys, st, graphs
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-sweep', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-graph', action='store_true')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
