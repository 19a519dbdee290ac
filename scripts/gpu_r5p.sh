#!/bin/bash
# Round 5, session p: builder A/B (warp-task threshold), config 3 with the scene block cache, tests touching the cache.
set -u
TAG=${1:-r5p}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_dynamic.py tests/test_gpu_replicate.py tests/test_gpu_build.py -x -q -m gpu 2>&1 | tail -4 | tee $OUT/${TAG}_pytest.txt
{
timeout 300 python scripts/build_ab.py
for LIB in rtbvh_b200/librtbvh_rs_w*.so; do RTBVH_LIB=$PWD/$LIB timeout 300 python scripts/build_ab.py; done
} 2>&1 | grep -v "^$" | tee $OUT/${TAG}_build_ab.txt
timeout 900 python bench.py --config 3 --no-cpu 2> $OUT/${TAG}_c3.err | tee $OUT/${TAG}_c3.json | cut -c1-300
tail -2 $OUT/${TAG}_c3.err
