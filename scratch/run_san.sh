#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 700 compute-sanitizer --tool memcheck python scratch/sanitize_smoke.py > gpurun_out/r2z_memcheck.log 2>&1
timeout 700 compute-sanitizer --tool synccheck python scratch/sanitize_smoke.py > gpurun_out/r2z_synccheck.log 2>&1
timeout 900 compute-sanitizer --tool racecheck python scratch/sanitize_smoke.py > gpurun_out/r2z_racecheck.log 2>&1
tail -3 gpurun_out/r2z_memcheck.log gpurun_out/r2z_synccheck.log gpurun_out/r2z_racecheck.log
