#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_tests.log 2>&1
tail -5 gpurun_out/r2b_tests.log
timeout 300 python scratch/graphprof.py > gpurun_out/r2b_graphprof.txt 2>&1
tail -14 gpurun_out/r2b_graphprof.txt
timeout 300 python scratch/l2probe.py > gpurun_out/r2b_l2probe.txt 2>&1
cat gpurun_out/r2b_l2probe.txt
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2b_bench.json').read().strip().splitlines()[-1])
print(round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['stages_ms'].items()})
for r in d.get('sweep',[]): print(round(r['ms_per_step'],4), r['case'])
PY
