#!/bin/bash
# Development aid: lateral sweep time per launch mode on the bench workload (run under gpurun).
# usage: mode_sweep.sh <basins> [modes...]   modes: grid cta cluster2 cluster4 cluster8 auto
B=${1:-2}; shift
MODES=${@:-grid cta cluster2 cluster4 cluster8 auto}
run() { python bench.py --basins $B --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$1', d['lateral_plan'], d['kernels_ms_per_step'])
    "; }
for m in $MODES; do
  case $m in
    grid) WF_LATERAL_MODE=grid run grid;;
    cta) WF_LATERAL_MODE=cta run cta;;
    cluster*) WF_LATERAL_MODE=cluster WF_CLUSTER_SIZE=${m#cluster} run $m;;
    auto) run auto;;
  esac
done
