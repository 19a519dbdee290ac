#!/bin/bash
set -u
TAG=${1:-r5v}
OUT=gpurun_out
mkdir -p $OUT
{
timeout 300 python scripts/build_ab.py
for LIB in rtbvh_b200/librtbvh_rs_*.so; do RTBVH_LIB=$PWD/$LIB timeout 300 python scripts/build_ab.py; done
} 2>&1 | grep -E "soup1m|soup4m" | grep sah | tee $OUT/${TAG}_build_ab.txt
