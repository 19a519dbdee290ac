#!/bin/bash
set -u
TAG=${1:-r5g}
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_replicate.py -x -q -m gpu 2>&1 | tail -5 | tee $OUT/${TAG}_pytest.txt
{
RTBVH_PACKET_MODE=lane timeout 300 python scripts/trace_ab.py --packets --name packet_lane 2>&1 | tail -1
for LIB in rtbvh_b200/librtbvh_rs_*.so; do
  RTBVH_LIB=$PWD/$LIB timeout 300 python scripts/trace_ab.py --packets 2>&1 | tail -1
  RTBVH_LIB=$PWD/$LIB timeout 300 python scripts/trace_ab.py --packets --any 2>&1 | tail -1
done
} | tee $OUT/${TAG}_ab.txt
