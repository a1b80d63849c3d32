#!/usr/bin/env python
"""Summarise an `ncu --set full` report: the metrics this repository quotes, per kernel launch.
usage: python profiles/summarize_full.py gpurun_out/r2_full.ncu-rep [traffic.json]
(reads the report with `ncu -i <rep> --page raw --csv`; with a second argument also rewrites the per-stage DRAM traffic
figures bench.py reports as `roofline.traffic`)."""
import csv
import io
import json
import subprocess
import sys

METRICS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct',
           'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
           'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active',
           'sm__warps_active.avg.pct_of_peak_sustained_active',
           'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
           'launch__grid_size', 'launch__block_size', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.max']
STAGE = {('tc_gram_kernel', '1'): 'wc_stats_fwd', ('tc_gram_kernel', '0'): 'wc_reduce_bwd'}


def main(rep, traffic_out=None):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, body = rows[0], rows[1], rows[2:]
    col = {n: i for i, n in enumerate(head)}
    seen, traffic, n_apply = {}, {}, 0
    for r in body:
        name = r[col['Kernel Name']]
        seen[name] = seen.get(name, 0) + 1
        print(f'{name[:100]}   (launch {seen[name]} of this kernel in the capture)')
        for m in METRICS:
            if m in col:
                print(f'    {m:100s} {r[col[m]]:>14s} {units[col[m]]}')
        print()
        def num(m):
            v = float(r[col[m]].replace(',', ''))
            u = units[col[m]].lower()
            return v * (1e9 if u.startswith('g') else 1e6 if u.startswith('m') else 1e3 if u.startswith('k') else 1)
        if 'dram__bytes_read.sum' in col:
            tot = num('dram__bytes_read.sum') + num('dram__bytes_write.sum')
            if 'tc_apply_kernel' in name:
                n_apply += 1
                traffic.setdefault('wc_apply_fwd' if n_apply % 2 == 1 else 'wc_apply_bwd', int(tot))   # K3 then K6 in stream order
            elif 'tc_gram_kernel' in name:
                traffic.setdefault('wc_stats_fwd' if ', 1>' in name or ', true>' in name else 'wc_reduce_bwd', int(tot))
    if traffic_out:
        traffic['source'] = ('profiles/r02_ncu_full_summary.txt: dram__bytes_read.sum + dram__bytes_write.sum per launch '
                             '(ncu --set full, cold L2; dirty lines still in L2 at kernel end are not counted as writes)')
        json.dump(traffic, open(traffic_out, 'w'), indent=1)


if __name__ == '__main__':
    main(*sys.argv[1:3])
