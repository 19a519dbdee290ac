#!/bin/bash
set -u
TAG=${1:-r6f}
OUT=gpurun_out
mkdir -p $OUT
{
timeout 300 python scripts/trace_ab.py --name default 2>&1 | tail -1
timeout 300 python scripts/trace_ab.py --packets --name default 2>&1 | tail -1
for LIB in rtbvh_b200/librtbvh_rs_*.so; do
  RTBVH_LIB=$PWD/$LIB timeout 300 python scripts/trace_ab.py 2>&1 | tail -1
  RTBVH_LIB=$PWD/$LIB timeout 300 python scripts/trace_ab.py --packets 2>&1 | tail -1
done
} | tee $OUT/${TAG}_ab.txt
