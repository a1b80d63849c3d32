#!/usr/bin/env python
"""bench.py -- WC layer fwd+bwd throughput on B200 (BASELINE.json's metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (weak scaling: one full shard per rank)
    python bench.py --gpus N --scaling strong ...            # fixed global batch 256x16x16x256 split over the ranks
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


def shard_shape(n_gpus: int, scaling: str):
    if scaling == 'strong':
        if SHAPE[0] % n_gpus:
            raise SystemExit(f'--scaling strong needs the batch {SHAPE[0]} divisible by --gpus {n_gpus}')
        return (SHAPE[0] // n_gpus,) + SHAPE[1:]
    return SHAPE


def workload_config(n_gpus: int, scaling: str = 'weak') -> dict:
    sh = shard_shape(n_gpus, scaling)
    return {'workload': 'cfg3: WC layer fwd+bwd on 256x16x16x256 fp32 NHWC per GPU' if scaling == 'weak' else
                        f'cfg3: WC layer fwd+bwd on a GLOBAL 256x16x16x256 fp32 NHWC batch, {sh[0]}x16x16x256 per GPU',
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
    steps = max(1, min(args.steps, 20))            # bounded sample: a CPU step is ~100 ms
    ts, cores = cpu_reference_time(SHAPE, steps, args.warmup)
    total = sum(ts)
    val = alg_bytes(SHAPE) * len(ts) / total / 1e9
    sample = f'{len(ts)} full fwd+bwd steps of the same 256x16x16x256 workload (one shard), torch-CPU fp32 port of the TF graph, {cores} threads'
    out = {'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': len(ts),
           'warmup': args.warmup, 'ms_per_step': 1e3 * total / len(ts), 'higher_is_better': True,
           'scaling': args.scaling, 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
           'config': workload_config(args.gpus, args.scaling),
           'per_step_ms': percentiles([1e3 * t for t in ts]),
           'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
           'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
           'note': 'TensorFlow 1.5 / Keras 2.0.8 cannot be installed here (no network, Python 3.12): the reference arm is '
                   'oracle/wc_oracle.py, the op-for-op CPU restatement of gan/conditional_layers.py:646-712 + generator.py:49-51'}
    print(json.dumps(out), flush=True)


def percentiles(v):
    v = sorted(v)
    q = lambda f: v[min(len(v) - 1, int(f * (len(v) - 1) + 0.5))]
    return {'median': q(0.5), 'p10': q(0.1), 'p90': q(0.9), 'min': v[0], 'max': v[-1], 'n': len(v)}


def numa_cpulists():
    """{node: set of CPUs} from sysfs, restricted to the CPUs this process may run on."""
    import glob
    allowed = os.sched_getaffinity(0)
    nodes = {}
    for d in glob.glob('/sys/devices/system/node/node[0-9]*'):
        cpus = set()
        for part in open(d + '/cpulist').read().strip().split(','):
            if part:
                a, _, b = part.partition('-')
                cpus.update(range(int(a), int(b or a) + 1))
        if cpus & allowed:
            nodes[int(d.rsplit('node', 1)[1])] = cpus & allowed
    return nodes


def probe_numa_node(torch, local: int):
    """sysfs does not say which NUMA node the GPU hangs off (containers report -1): measure it.  For every node, run on its
    CPUs, pin 64 MB of freshly touched host memory there (first-touch placement) and time host -> device copies; stay on
    the fastest node.  Returns the node, or None when there is nothing to choose."""
    nodes = numa_cpulists()
    if len(nodes) < 2:
        return None
    before = os.sched_getaffinity(0)
    dev = torch.device('cuda', local)
    dst = torch.empty(16 << 20, dtype=torch.float32, device=dev)
    best, best_ms, seen = None, None, {}
    for node, cpus in sorted(nodes.items()):
        try:
            os.sched_setaffinity(0, cpus)
            h = torch.empty(16 << 20, dtype=torch.float32).fill_(1.0).pin_memory()
            for _ in range(2):
                dst.copy_(h, non_blocking=True)
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(4):
                dst.copy_(h, non_blocking=True)
            e1.record()
            torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1) / 4
            seen[node] = round(64 * 1.048576 / ms, 1)         # GB/s
            if best_ms is None or ms < best_ms:
                best, best_ms = node, ms
            del h
        except Exception:
            continue
    if best is None or max(seen.values()) < 1.1 * min(seen.values()):
        os.sched_setaffinity(0, before)                       # no real difference: leave the placement alone
        return None
    os.sched_setaffinity(0, nodes[best])
    sys.stderr.write(f'[bench] host -> device GB/s by NUMA node {seen}: staying on node {best}\n')
    return best


def pin_to_gpu_numa_node(torch, local: int):
    """Bind this rank (and so the first-touch placement of its pinned host buffers) to the NUMA node of its GPU: the e2e
    leg is a PCIe stream, and with every rank on node 0 the ranks behind the other socket cross the inter-socket link."""
    try:
        bdf = torch.cuda.get_device_properties(local).pci_bus_id if hasattr(torch.cuda.get_device_properties(local), 'pci_bus_id') else None
        if bdf is None:
            out = subprocess.run(['nvidia-smi', '--query-gpu=pci.bus_id', '--format=csv,noheader', '-i', str(local)],
                                 capture_output=True, text=True).stdout.strip()
            bdf = out[-12:] if len(out) >= 12 else out
        bdf = str(bdf).lower()
        if len(bdf.split(':')[0]) == 8:
            bdf = bdf[4:]
        node = int(open(f'/sys/bus/pci/devices/{bdf}/numa_node').read().strip())
        if node < 0:
            return probe_numa_node(torch, local)
        cpus = set()
        for part in open(f'/sys/devices/system/node/node{node}/cpulist').read().strip().split(','):
            a, _, b = part.partition('-')
            cpus.update(range(int(a), int(b or a) + 1))
        os.sched_setaffinity(0, cpus & os.sched_getaffinity(0) or cpus)
        return node
    except Exception:
        return None


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
    numa = pin_to_gpu_numa_node(torch, local)
    LSHAPE = shard_shape(world, args.scaling)          # this rank's shard
    # rotate over enough buffer sets that one step's inputs are not L2-resident from the previous one (> 2x L2 touched)
    NBUF = max(4, min(32, -(-268 * 2 ** 20 // (2 * 4 * LSHAPE[0] * LSHAPE[1] * LSHAPE[2] * LSHAPE[3]))))
    xs, dys = make_inputs(torch, LSHAPE, NBUF, dev, seed=1234 + rank)
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
    value = world * alg_bytes(LSHAPE) / (per_step * 1e-3) / 1e9
    hbm, peak_src = peaks()

    # ---- distribution of single steps: 200 more replays, one CUDA event pair per step (median,
    #      p10 / p90; the headline above stays the contract's K-step mean)
    nper = 200
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(nper + 1)]
    if barrier:
        barrier()
    torch.cuda.synchronize()
    evs[0].record()
    for i in range(nper):
        step(i)
        evs[i + 1].record()
    torch.cuda.synchronize()
    per_step_ms = percentiles([evs[i].elapsed_time(evs[i + 1]) for i in range(nper)])

    # ---- per-stage device time: the four M x C stages launched standalone through the fine-grained C ABI on the
    #      same rotating buffers, CUDA events on the launching stream around `n` back-to-back launches
    stages = stage_times(torch, lib, xs, dys, min(max(args.steps, 10), 50), LSHAPE)
    M, C = LSHAPE[0] * LSHAPE[1] * LSHAPE[2], LSHAPE[3]
    stage_bytes = {'wc_stats_fwd': 4 * M * C, 'wc_apply_fwd': 8 * M * C, 'wc_reduce_bwd': 8 * M * C, 'wc_apply_bwd': 12 * M * C}
    dom = max(stage_bytes, key=lambda k: stages.get(k, 0.0))
    ach = stage_bytes[dom] / (stages[dom] * 1e-3) / 1e9
    # traffic: dram__bytes_read.sum + dram__bytes_write.sum of that kernel from the committed `ncu --set full` capture of the
    # full-size shard (a profiler pass cannot run inside a timed bench): a STATIC figure, null for any other shard size
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tp) and LSHAPE == SHAPE:
        try:
            tj = json.load(open(tp))
            traffic, traffic_src = tj.get(dom), 'static: ' + str(tj.get('source', 'profiles/traffic.json'))
        except Exception:
            traffic = None
    roofline = {'bound': 'hbm', 'kernel': dom, 'achieved': ach, 'peak': hbm, 'unit': 'GB/s', 'frac': ach / hbm,
                'traffic': traffic, 'traffic_source': traffic_src, 'peak_source': peak_src,
                'algorithmic_bytes_per_launch': stage_bytes[dom], 'avg_launch_ms': stages[dom],
                'all_stages_frac': {k: stage_bytes[k] / (stages[k] * 1e-3) / 1e9 / hbm for k in stage_bytes}}
    # The same kernels against the OTHER roof: fp32-accurate products on bf16 tensor cores are three bf16 products each
    # (hi.hi + hi.lo + lo.hi), so the path issues 6 M C^2 (K1: 3 of 4 quadrants -> 4.5), 6 (K3), 6 (K4), 12 (K6) M C^2 bf16
    # flops.  At C = 256 that is as much tensor time as HBM time: reported beside the HBM fraction, not instead of it.
    try:
        tf_peak = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['bf16_tflops'])
    except Exception:
        tf_peak = 2250.0
    stage_flops = {'wc_stats_fwd': 4.5 * M * C * C, 'wc_apply_fwd': 6.0 * M * C * C, 'wc_reduce_bwd': 6.0 * M * C * C,
                   'wc_apply_bwd': 12.0 * M * C * C}
    roofline['tensor'] = {'unit': 'TFLOP/s', 'peak': tf_peak, 'what': 'bf16 flops actually issued (3 bf16 products per fp32-accurate product) / stage time',
                          'flops_per_launch': stage_flops[dom], 'achieved': stage_flops[dom] / (stages[dom] * 1e-3) / 1e12,
                          'frac': stage_flops[dom] / (stages[dom] * 1e-3) / 1e12 / tf_peak,
                          'all_stages_frac': {k: stage_flops[k] / (stages[k] * 1e-3) / 1e12 / tf_peak for k in stage_flops},
                          'min_ms_at_peak': {k: stage_flops[k] / tf_peak / 1e9 for k in stage_flops},
                          'min_ms_at_hbm_peak': {k: stage_bytes[k] / hbm / 1e6 for k in stage_bytes}}

    out = None
    if rank == 0:
        out = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
               'ms_per_step': per_step, 'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None, 'dtype': 'f32',
               'data': 'synthetic', 'config': workload_config(world, args.scaling), 'roofline': roofline,
               'per_step_ms': per_step_ms, 'buffer_sets': NBUF, 'numa_node': numa,
               'hbm_frac_fwd_bwd': value / world / hbm, 'hbm_frac_fwd_bwd_single_touch': value / world / hbm / 2,
               'stages_ms': stages, 'gpu_launches': int(launches), 'clocks': clocks,
               'path': int(lib.wc_path_for(M, C, C, C)), 'cuda_graph': bool(graphs)}

    # ---- end to end through the public module with HOST buffers (H2D + D2H inside the timed region)
    # Every step copies ITS inputs from pinned host memory and ITS results back; the three legs run on three streams
    # with two buffer sets, so step i's read-back overlaps step i+1's upload (PCIe is full duplex) and both overlap
    # the kernels of the step in between.
    xh = [x.cpu().pin_memory() for x in xs[:2]]
    dyh = [d.cpu().pin_memory() for d in dys[:2]]
    yh = [torch.empty(LSHAPE, dtype=torch.float32).pin_memory() for _ in range(2)]
    dxh = [torch.empty(LSHAPE, dtype=torch.float32).pin_memory() for _ in range(2)]
    xd = [torch.empty(LSHAPE, device=dev) for _ in range(2)]
    dyd = [torch.empty(LSHAPE, device=dev) for _ in range(2)]
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
        # The results of the step that used this slot two steps ago are released HERE, on the compute stream, after that
        # stream has been made to wait for their read-back: the caching allocator can then hand the same blocks to this
        # step.  (With Tensor.record_stream instead, a block becomes reusable only once an event recorded at free time
        # has completed; the host runs several steps ahead of the device, so every step allocated fresh 67 MB blocks
        # with cudaMalloc -- hidden on boxes where that call is fast, 15 ms per step where it is not.)
        cur.wait_event(ev_out[b])
        alive[b] = None
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
        alive[b] = (y, gr)

    def e2e_drain():
        cur.wait_stream(s_in); cur.wait_stream(s_out)

    # What the host link of THIS box delivers for the same buffers, measured right here (boxes differ by 5x): the pure
    # copies of one step (x, dy up; y, dx down) one direction at a time and both at once.  The e2e step cannot be
    # faster than the duplex figure; e2e['link'] lets a reader see how close it is.
    def link_probe():
        yd0, dxd0 = torch.empty(LSHAPE, device=dev), torch.empty(LSHAPE, device=dev)
        def up():
            with torch.cuda.stream(s_in):
                xd[0].copy_(xh[0], non_blocking=True); dyd[0].copy_(dyh[0], non_blocking=True)
        def down():
            with torch.cuda.stream(s_out):
                yh[0].copy_(yd0, non_blocking=True); dxh[0].copy_(dxd0, non_blocking=True)
        res = {}
        for name, legs in (('h2d', (up,)), ('d2h', (down,)), ('duplex', (up, down))):
            for leg in legs:
                leg()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(3):
                for leg in legs:
                    leg()
            torch.cuda.synchronize()
            res[name + '_ms_per_step'] = (time.perf_counter() - t0) / 3 * 1e3
        return res
    link = link_probe()

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
        out['e2e'] = {'value': world * alg_bytes(LSHAPE) / (ems / e_steps * 1e-3) / 1e9, 'unit': UNIT,
                      'h2d_bytes_per_step': 2 * nbytes, 'd2h_bytes_per_step': 2 * nbytes, 'ms_per_step': ems / e_steps,
                      'steps': e_steps,
                      'link': dict(link, note='pure pinned-memory copies of one step\'s 134 MB each way on this box, same '
                                              'buffers, measured just before the e2e loop: the step cannot beat duplex_ms_per_step'),
                      'link_frac': link['duplex_ms_per_step'] / (ems / e_steps),
                      'what': 'every step: pinned host x, dy -> device; module fwd+bwd; y, dx -> pinned host; upload, '
                              'kernels and read-back of consecutive steps overlap on three streams (two buffer sets)'}

    # ---- parity of THIS launch configuration against the fp64 oracle on the concatenated batch (small shards, so that the
    #      oracle takes seconds): with N > 1 it is the evidence that synchronised-batch whitening equals the single-process layer
    par = parity_block(torch, dist, W, dev, rank, world, group)
    if rank == 0:
        out['parity'] = par

    # ---- other configurations (parity-test cases; reported, not the headline)
    if rank == 0 and world == 1 and not args.no_sweep:
        del xs[1:], dys[1:]
        out['sweep'] = sweep(torch, W, dev, hbm)
        out['layer_lists'] = layer_lists(torch, W, dev, hbm)
        out['gan_step'] = gan_step(torch, W, dev)

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


def parity_block(torch, dist, W, dev, rank, world, group):
    """Every rank runs a 16x8x8x256 shard of one seeded global batch through create_norm('d','ufconv') (K = 5, E = 3) with
    the launch's process group; rank 0 gathers inputs, outputs and gradients and compares them with the fp64 oracle on
    the concatenated batch.  The oracle is used here as the checker only."""
    from oracle import wc_oracle as O
    n_loc, K, E, C = 16, 5, 3, 256
    shape = (n_loc * world, 8, 8, C)
    x_all = O.make_activation(shape, 'correlated', seed=11)
    gen = torch.Generator().manual_seed(12)
    cls_all = torch.randint(0, K, (shape[0], 1), generator=gen, dtype=torch.int32)
    dy_all = torch.randn(shape, generator=gen)
    params = O.make_params(C, after_norm='ufconv', number_of_classes=K, filters_emb=E, seed=13, random_bias=True)
    stack = W.create_norm('d', 'ufconv', number_of_classes=K, filters_emb=E, process_group=group)(axis=-1, name='parity')
    stack.build(C, dev); stack.built = True
    stack.to(dev).train()
    with torch.no_grad():
        stack.repart_c.class_matrix.copy_(params['class_matrix']); stack.repart_c.kernel.copy_(params['f_kernel'])
        stack.repart_u.kernel.copy_(params['u_kernel']); stack.repart_u.bias.copy_(params['u_bias'])
    sl = slice(rank * n_loc, (rank + 1) * n_loc)
    x = x_all[sl].to(dev).requires_grad_(True)
    y = stack(x, cls_all[sl].to(dev))
    y.backward(dy_all[sl].to(dev))
    torch.cuda.synchronize()
    stack.check_info()
    # coloring-parameter gradients are per-shard sums (DDP semantics): add them up for the comparison
    pg = torch.cat([p.grad.reshape(-1) for p in (stack.repart_u.kernel, stack.repart_u.bias, stack.repart_c.kernel, stack.repart_c.class_matrix)])
    mv = stack.npart.moving_variance.detach().clone()
    if world > 1:
        ys = [torch.empty_like(y) for _ in range(world)]; dxs = [torch.empty_like(x.grad) for _ in range(world)]
        mvs = [torch.empty_like(mv) for _ in range(world)]
        dist.all_gather(ys, y.detach().contiguous()); dist.all_gather(dxs, x.grad.contiguous()); dist.all_gather(mvs, mv)
        dist.all_reduce(pg)
        y_g, dx_g = torch.cat(ys), torch.cat(dxs)
        same_stats = all(bool(torch.equal(mvs[0], m)) for m in mvs)
    else:
        y_g, dx_g, same_stats = y.detach(), x.grad, True
    if rank != 0:
        return None
    p64 = {k: v.double() for k, v in params.items()}
    leaves = {k: p64[k].clone().requires_grad_(True) for k in ('u_kernel', 'u_bias', 'f_kernel', 'class_matrix')}
    p64.update(leaves)
    x64 = x_all.double().requires_grad_(True)
    yr, _, mcr = O.wc_block(x64, cls_all, p64, norm='d', after_norm='ufconv', training=True)
    gr = torch.autograd.grad(yr, [x64] + list(leaves.values()), dy_all.double())
    rel = lambda a, b: float((a.detach().double().cpu() - b.detach()).abs().max() / b.detach().abs().max())
    pr = torch.cat([g.reshape(-1) for g in gr[1:]])
    from wc_gan_b200 import sync
    transport = 'single process' if world == 1 else ('peer memory (in-kernel exchange)' if sync.peer_exchange(group, C, C, dev) is not None else 'nccl collectives')
    return {'case': f"create_norm('d','ufconv') K={K} E={E}, global batch {shape[0]}x8x8x{C} = {world} shard(s) of {n_loc}",
            'oracle': 'oracle/wc_oracle.py fp64 on the concatenated batch', 'transport': transport,
            'y_rel_err': rel(y_g, yr), 'dx_rel_err': rel(dx_g, gr[0]), 'param_grad_rel_err': rel(pg, pr),
            'moving_cov_rel_err': rel(mv, mcr), 'moving_stats_identical_on_all_ranks': same_stats,
            'tolerance': {'outputs': 1e-4, 'gradients': 1e-3}}


def stage_times(torch, lib, xs, dys, n, shape=None):
    """ms per launch of K1 (wc_stats_fwd), K3 (wc_apply_fwd), K4 (wc_reduce_bwd), K6 (wc_apply_bwd) at the bench shape.
    Each stage includes its helper launches (memset, finalize).  Operand values do not matter for the timing."""
    from wc_gan_b200._lib import ptr
    N, H, Wd, C = shape or SHAPE
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


def time_case(torch, W, dev, shape, after_norm, K=None, E=10, *, fwd_only=False, relu=False, steps=20, nbuf=None):
    """ms per fwd(+bwd) of create_norm('d', after_norm[, activation='relu']) on `shape`, replayed from CUDA graphs over
    rotating buffer sets (> 2x L2 touched per round when the tensor is large; small tensors are L2-resident by size)."""
    nbytes = 4 * shape[0] * shape[1] * shape[2] * shape[3]
    nbuf = nbuf or max(2, min(8, -(-268 * 2 ** 20 // (2 * nbytes))))
    xs, dys = make_inputs(torch, shape, nbuf, dev, seed=7)
    st = W.create_norm('d', after_norm, number_of_classes=K, filters_emb=E, activation='relu' if relu else None)(axis=-1, name='b')
    st.build(shape[-1], dev); st.built = True
    st.to(dev).train()
    params = [p for p in st.parameters()]
    cls = torch.randint(0, K, (shape[0], 1), dtype=torch.int32, device=dev) if K else None

    def step(i):
        if fwd_only:
            with torch.no_grad():
                return st(xs[i % nbuf], cls)
        x = xs[i % nbuf].requires_grad_(True)
        y = st(x, cls)
        return torch.autograd.grad(y, [x] + params, dys[i % nbuf])
    graphs = []
    for i in range(3):
        step(i)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for i in range(nbuf):
            step(i)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    keep = []
    for i in range(nbuf):
        gph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gph, stream=side):
            keep.append(step(i))
        graphs.append(gph)
    ms, _, _ = time_steps(torch, lambda i: graphs[i % nbuf].replay(), steps, 3)
    del graphs, keep
    return ms / steps


def layer_lists(torch, W, dev, hbm):
    """The WC layers wc-gan actually runs (BASELINE.json configs[3], configs[4]; SURVEY.md section 8 "cfg4" / "cfg5"): the
    generator's layer list timed layer by layer through create_norm with the block's ReLU fused, at the generator-step batch
    (fwd+bwd, 2B: gan/train.py:85-91) and at the discriminator-step batch (generator forward only, B)."""
    out = []
    for name, B, K, E, res in (('cfg4 cifar10_resnet_sn_cond_sa generator (7 WC layers, C=128, K=10, E=4)', 256, 10, 4, (4, 8, 8, 16, 16, 32)),
                               ('cfg5 tinyimagenet_resnet_sn_cond_sa generator (9 WC layers, C=128, K=200, E=15), 1024 / 8 per GPU', 128, 200, 15,
                                (4, 8, 8, 16, 16, 32, 32, 64))):
        layers = [(r, 'ufconv') for r in res] + [(res[-1], 'uconv')]
        memo, rows = {}, []
        tot_g = tot_d = byt_g = 0.0
        for r, an in layers:
            key = (r, an)
            if key not in memo:
                k = K if an == 'ufconv' else None
                memo[key] = (time_case(torch, W, dev, (2 * B, r, r, 128), an, k, E, relu=True),
                             time_case(torch, W, dev, (B, r, r, 128), an, k, E, relu=True, fwd_only=True))
            g_ms, d_ms = memo[key]
            rows.append({'layer': f'{r}x{r}x128 {an}', 'g_step_fwd_bwd_ms': g_ms, 'd_step_fwd_ms': d_ms,
                         'g_step_hbm_frac': alg_bytes((2 * B, r, r, 128)) / (g_ms * 1e-3) / 1e9 / hbm})
            tot_g += g_ms; tot_d += d_ms; byt_g += alg_bytes((2 * B, r, r, 128))
        out.append({'list': name, 'batch': B, 'layers': rows, 'g_step_wc_ms': tot_g, 'd_step_wc_ms': tot_d,
                    'wc_ms_per_iteration': 5 * tot_d + tot_g, 'g_step_gbs': byt_g / (tot_g * 1e-3) / 1e9,
                    'g_step_hbm_frac': byt_g / (tot_g * 1e-3) / 1e9 / hbm,
                    'iteration': '5 discriminator steps (generator forward at B) + 1 generator step (fwd+bwd at 2B), gan/train.py:78-99'})
    return out


def gan_step(torch, W, dev):
    """BASELINE.json configs[3]: one full cifar10_resnet_sn_cond_sa training iteration (generator with WC layers + projection
    discriminator with spectral norm, hinge loss, Adam(0, 0.9), batch 256, 5 D steps + 1 G step at 2B) on synthetic 32x32
    images.  Convolutions are cuDNN via PyTorch (out of the hot path); the WC layers are libwc_b200."""
    import time as _t
    torch.manual_seed(0)
    G = W.make_generator(number_of_classes=10, block_norm='d', block_after_norm='ufconv', filters_emb=4, last_norm='d',
                         last_after_norm='uconv').to(dev)
    D = W.make_discriminator((32, 32, 3), block_sizes=(256,) * 4, number_of_classes=10, type='PROJECTIVE', spectral=True,
                             sum_pool=True).to(dev)
    tr = W.GanTrainer(G, D, batch_size=256, training_ratio=5, generator_batch_multiple=2, total_generator_iters=50000.0, device=dev)
    for _ in range(3):
        tr.train_one_step()
    torch.cuda.synchronize()
    n = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        tr.train_one_step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    for s_ in G.wc_layers():
        s_.check_info()
    return {'config': 'cfg4: cifar10_resnet_sn_cond_sa, batch 256, training_ratio 5, generator_batch_multiple 2, synthetic 32x32x3',
            'ms_per_iteration': ms, 'iters_per_s': 1e3 / ms, 'mode': 'eager PyTorch (cuDNN convolutions) + libwc_b200 WC layers, 1 GPU',
            'note': 'the reference reports no iters/s; its hardware was a single GPU of 2018 (BASELINE.md)'}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'],
                    help='weak: every rank holds its own 256x16x16x256 shard (default); strong: one global 256x16x16x256 batch split over the ranks')
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
