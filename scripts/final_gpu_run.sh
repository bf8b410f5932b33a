#!/bin/bash
# Round-end measurement on one B200 (run under gpurun): GPU tests, the default bench line, the ncu launch list of the
# bench command and one full ncu capture of the two hot kernels.  Outputs under gpurun_out/.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/gpu_tests.log
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; cat gpurun_out/bench_default.json
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-single-basin"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on -k 'regex:k_vertical|k_lateral_wave' --launch-skip 4 --launch-count 2 \
  -f -o gpurun_out/prof_r1_final $CMD > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/prof_r1_final.ncu-rep gpurun_out/launches_r1.csv
