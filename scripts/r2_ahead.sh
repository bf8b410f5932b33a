#!/bin/bash
# Development aid (run under gpurun): the build variants of the staged sweep on the bench shard (checksums must agree),
# then the whole 16384^2 grid (single-block sweeps, eight items per thread and tick) with the default build.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python scripts/variant_bench.py compare 32 ${VARIANTS:-base ahead aheadtp aheadinl rolled} 2>&1 | tee gpurun_out/variants_ahead.log
if [ -n "${FULLLIB:-}" ]; then
  WFLOW_B200_LIB=$PWD/wflow_b200/libwflowsbm_$FULLLIB.so timeout 600 python bench.py --config 3 --basins 256 --steps 5 --warmup 3 --no-cpu-baseline --no-single-basin > gpurun_out/bench_full16k_$FULLLIB.json 2> gpurun_out/bench_full16k_$FULLLIB.err
  python -c "import json,sys; d=json.loads(open('gpurun_out/bench_full16k_$FULLLIB.json').read().strip().splitlines()[-1]); print('full16k $FULLLIB', d['ms_per_step'], d['kernels_ms_per_step'])"
fi
