#!/bin/bash
# Development aid: lateral sweep time per launch mode on the bench workload (run under gpurun).
B=${1:-2}
run() { python bench.py --basins $B --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$1', d['lateral_plan'], d['kernels_ms_per_step'])
    "; }
WF_LATERAL_MODE=grid run grid
WF_LATERAL_MODE=cta run cta
for cs in 2 4 8; do WF_LATERAL_MODE=cluster WF_CLUSTER_SIZE=$cs run cluster$cs; done
run auto
