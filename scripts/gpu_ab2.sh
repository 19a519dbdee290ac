#!/bin/bash
# A/B of every librtbvh_rs_<variant>.so plus the trace modes of the default library with scripts/trace_ab.py (seconds each).
set -u
TAG=${1:-ab2}
OUT=gpurun_out
mkdir -p $OUT
{
for MODE in persistent phased; do
  RTBVH_TRACE_MODE=$MODE timeout 300 python scripts/trace_ab.py --name default 2>&1 | tail -1
done
for LIB in rtbvh_b200/librtbvh_rs_*.so; do
  RTBVH_LIB=$PWD/$LIB timeout 300 python scripts/trace_ab.py 2>&1 | tail -1
done
} | tee $OUT/${TAG}_ab.txt
