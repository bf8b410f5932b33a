#!/bin/bash
# Development aid (run under gpurun): full GPU tests, per-task contributions to the lateral sweep, per-tick profile.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/gpu_tests_full.log; tail -5 gpurun_out/gpu_tests_full.log
timeout 600 python scripts/lat_bench.py 32 WF_LATERAL_TASKS=7,1,2,4,3,5,6 > gpurun_out/lat_tasks.log 2>&1; tail -16 gpurun_out/lat_tasks.log
WFLOW_B200_LIB=wflow_b200/libwflowsbm_prof.so WF_TICK_PROFILE=gpurun_out/ticks_r2.csv timeout 600 python scripts/tick_run.py 32 > gpurun_out/tick_run.log 2>&1
python scripts/tick_profile.py gpurun_out/ticks_r2.csv | tee gpurun_out/tick_profile_r2.txt
