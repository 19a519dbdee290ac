#!/bin/bash
set -u
TAG=${1:-r5w}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_traversal.py -x -q -m gpu -k "oddly or edge or teapot" 2>&1 | tail -3
{
timeout 300 python scripts/build_ab.py
for LIB in rtbvh_b200/librtbvh_rs_*.so; do RTBVH_LIB=$PWD/$LIB timeout 300 python scripts/build_ab.py; done
} 2>&1 | grep -E "sah" | tee $OUT/${TAG}_build_ab.txt
