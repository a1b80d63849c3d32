#!/bin/bash
# round-2 baseline: tests, L2-hint A/B, graph timeline, sanitizer logs
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1
for l2 in 0 1 5 7 15 63 56 21; do
  WC_B200_L2=$l2 timeout 300 python bench.py --steps 100 --warmup 5 --no-sweep --no-cpu > gpurun_out/r2a_bench_l2_$l2.json 2> gpurun_out/r2a_bench_l2_$l2.err
done
timeout 300 python scratch/graphprof.py > gpurun_out/r2a_graphprof.txt 2>&1
timeout 500 compute-sanitizer --tool memcheck python scratch/sanitize_smoke.py > gpurun_out/r2a_memcheck.log 2>&1
timeout 500 compute-sanitizer --tool synccheck python scratch/sanitize_smoke.py > gpurun_out/r2a_synccheck.log 2>&1
timeout 500 compute-sanitizer --tool racecheck python scratch/sanitize_smoke.py > gpurun_out/r2a_racecheck.log 2>&1
tail -5 gpurun_out/r2a_tests.log
for l2 in 0 1 5 7 15 63 56 21; do python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2a_bench_l2_$l2.json').read().strip().splitlines()[-1])
    print('L2=$l2', round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['stages_ms'].items()})
except Exception as e: print('L2=$l2 failed', e)
PY
done
tail -3 gpurun_out/r2a_memcheck.log gpurun_out/r2a_synccheck.log gpurun_out/r2a_racecheck.log
