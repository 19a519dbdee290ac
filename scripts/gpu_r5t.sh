#!/bin/bash
set -u
TAG=${1:-r5t}
OUT=gpurun_out
mkdir -p $OUT
echo "== packet parity with the lane kernels forced for both trees"
RTBVH_PACKET_MODE=lane timeout 900 python -m pytest tests/test_gpu_traversal.py tests/test_zz_gpu_golden.py tests/test_gpu_dynamic.py -x -q -m gpu 2>&1 | tail -4 | tee $OUT/${TAG}_pytest_lane.txt
{
for M in static lane; do
  RTBVH_PACKET_MODE=$M timeout 300 python scripts/trace_ab.py --packets --bvh --name bvh_packet_$M 2>&1 | tail -1
  RTBVH_PACKET_MODE=$M timeout 300 python scripts/trace_ab.py --packets --bvh --any --name bvh_packet_$M 2>&1 | tail -1
done
} | tee $OUT/${TAG}_ab.txt
for V in default nostream; do
  if [ $V = nostream ]; then export RTBVH_LIB=$PWD/rtbvh_b200/librtbvh_rs_nostream.so; fi
  timeout 900 python bench.py --config 4 --no-cpu --steps 10 --warmup 3 --e2e-steps 2 2> $OUT/${TAG}_c4_$V.err > $OUT/${TAG}_c4_$V.json
  python -c "
import json; d=json.load(open('$OUT/${TAG}_c4_$V.json')); print('config 4 $V', round(d['value'],1), 'e2e', round(d['e2e']['value'],1))"
  timeout 900 python bench.py --config 5 --no-cpu --steps 10 --warmup 3 --e2e-steps 2 2> $OUT/${TAG}_c5_$V.err > $OUT/${TAG}_c5_$V.json
  python -c "
import json; d=json.load(open('$OUT/${TAG}_c5_$V.json')); print('config 5 $V', round(d['value'],1), 'e2e', round(d['e2e']['value'],1))"
done
