#!/bin/bash
# final artefacts of the round: tests, bench line, graph timelines (cfg3, cfg2), ncu launch list and full capture
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/final_tests.log 2>&1; tail -2 gpurun_out/final_tests.log
timeout 600 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
timeout 120 python scratch/graphprof.py 2>&1 | grep -v -i warn > gpurun_out/final_graphprof_cfg3.txt
WC_SHAPE=64x16x16x128 WC_AN=ufconv WC_K=10 WC_E=4 timeout 120 python scratch/graphprof.py 2>&1 | grep -v -i warn > gpurun_out/final_graphprof_cfg2.txt
WC_SHAPE=64x8x8x256 timeout 120 python scratch/graphprof.py 2>&1 | grep -v -i warn > gpurun_out/final_graphprof_cfg1.txt
timeout 120 python scratch/chaintime.py 2>&1 | grep -v -i warn > gpurun_out/final_chaintime.txt
export WC_B200_NO_PDL=1
K='regex:tc_apply|tc_gram|chol_cluster|chain_kernel|stats_finalize|zero_kernel'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 30 -c 100 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-sweep --no-cpu --no-graph > gpurun_out/r2_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "$K" -s 40 -c 10 -o gpurun_out/r2_full -f python bench.py --steps 2 --warmup 3 --no-sweep --no-cpu --no-graph > gpurun_out/r2_ncu_full.log 2>&1
ls -la gpurun_out/r2_full.ncu-rep gpurun_out/r2_launches.csv; tail -2 gpurun_out/r2_ncu_full.log
