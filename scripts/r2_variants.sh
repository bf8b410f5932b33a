#!/bin/bash
# Development aid (run under gpurun): GPU tests, then the build variants of both kernels on the bench shard.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
if [ "${SKIP_TESTS:-0}" != 1 ]; then
  timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/gpu_tests_full.log; tail -4 gpurun_out/gpu_tests_full.log
fi
timeout 1200 python scripts/variant_bench.py compare 32 $VARIANTS 2>&1 | tee gpurun_out/variants.log
if [ -n "${VARIANTS2:-}" ]; then
  timeout 900 python scripts/variant_bench.py compare 32 $VARIANTS2 -- $ENV2 2>&1 | tee -a gpurun_out/variants.log
fi
