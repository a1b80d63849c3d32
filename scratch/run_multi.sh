#!/bin/bash
# N-GPU checks: 2-GPU parity tests, weak and strong scaling bench lines
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2g_multi_tests.log 2>&1; tail -3 gpurun_out/r2g_multi_tests.log
for sc in weak strong; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 5 --scaling $sc > gpurun_out/r2g_bench_${N}gpu_$sc.json 2> gpurun_out/r2g_bench_${N}gpu_$sc.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2g_bench_${N}gpu_$sc.json").read().strip().splitlines()[-1])
    print("$sc N=$N", round(d["ms_per_step"],4), round(d["value"],1), d["per_step_ms"]["median"], d["parity"], d["e2e"]["value"])
except Exception as e:
    print("$sc failed", e); print(open("gpurun_out/r2g_bench_${N}gpu_$sc.err").read()[-1500:])
PY
done
