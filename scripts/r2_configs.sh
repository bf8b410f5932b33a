#!/bin/bash
# Round-2 measurement of the other BASELINE configurations on one B200 (run under gpurun): config 2 (4096^2 column only,
# 1000 daily steps), config 4 (256 Rhine trials as one device model), config 3 as the WHOLE 16384^2 grid on one GPU.
# One JSON line each under gpurun_out/ (copied to profiles/r2_bench_*.json afterwards).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; free -g | head -2
run() {  # name timeout args...
  local name=$1 to=$2; shift 2
  local t0=$SECONDS
  timeout "$to" python bench.py "$@" > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  echo "== $name rc=$? $((SECONDS - t0)) s"; tail -c 1500 gpurun_out/bench_$name.json; echo; tail -3 gpurun_out/bench_$name.err
}
for c in ${CONFIGS:-2 4 full}; do
  case $c in
    2) run config2 600 --config 2 --steps 1000 --warmup 5 ;;
    4) run config4 600 --config 4 --steps 29 --warmup 3 ;;
    full) run config3_full16k 1200 --config 3 --basins 256 --steps 10 --warmup 3 --no-cpu-baseline --no-single-basin ;;
    5) run config5_n1 900 --config 5 --basins ${BASINS5:-32} --steps 20 --warmup 3 --no-cpu-baseline ;;
  esac
done
