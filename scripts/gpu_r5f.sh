#!/bin/bash
# Round 5, session f: parity of the new packet kernel + scene replication tests, then packet A/B.
set -u
TAG=${1:-r5f}
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_traversal.py tests/test_gpu_replicate.py tests/test_zz_gpu_golden.py tests/test_gpu_dynamic.py -x -q -m gpu 2>&1 | tail -15 | tee $OUT/${TAG}_pytest.txt
{
for M in lane persistent; do
  for F in "" "--any"; do
    RTBVH_PACKET_MODE=$M timeout 300 python scripts/trace_ab.py --packets $F --name packet_$M 2>&1 | tail -1
  done
done
for LIB in rtbvh_b200/librtbvh_rs_*.so; do
  RTBVH_LIB=$PWD/$LIB timeout 300 python scripts/trace_ab.py --packets 2>&1 | tail -1
done
timeout 300 python scripts/trace_ab.py --name single 2>&1 | tail -1
} | tee $OUT/${TAG}_ab.txt
