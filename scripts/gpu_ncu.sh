#!/bin/bash
# One full ncu capture (with source) of the traversal kernel in the bench workload.
set -u
TAG=${1:-ncu}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_single -s 3 -c 1 -f -o $OUT/${TAG}_prof \
    python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > $OUT/${TAG}_ncu_full.log 2>&1
tail -2 $OUT/${TAG}_ncu_full.log | cut -c1-200
