#!/bin/bash
# Round-2 measurement on one B200 (run under gpurun): GPU tests, the default bench line, the ncu launch list of the
# bench command and one full ncu capture of the two hot kernels.  Outputs under gpurun_out/ (tag = $1, default r2).
set -u
cd "$(dirname "$0")/.."
TAG=${1:-r2}
mkdir -p gpurun_out
if [ "${SKIP_TESTS:-0}" != 1 ]; then
  timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/gpu_tests_$TAG.log
fi
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; cat gpurun_out/bench_$TAG.json
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-single-basin"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_vertical|k_lateral_wave' --launch-skip 4 --launch-count 2 \
  -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/prof_$TAG.ncu-rep gpurun_out/launches_$TAG.csv
