#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
export WC_B200_NO_PDL=1
K='regex:tc_apply|tc_gram|chol_cluster|chain_kernel|stats_finalize|zero_kernel'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 30 -c 100 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-sweep --no-cpu --no-graph > gpurun_out/r2_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "$K" -s 40 -c 10 -o gpurun_out/r2_full -f python bench.py --steps 2 --warmup 3 --no-sweep --no-cpu --no-graph > gpurun_out/r2_ncu_full.log 2>&1
ls -la gpurun_out/r2_full.ncu-rep gpurun_out/r2_launches.csv; tail -2 gpurun_out/r2_ncu_full.log
