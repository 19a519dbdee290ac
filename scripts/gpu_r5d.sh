#!/bin/bash
# Round 5, session d: configs 4 and 5 in both refill-kernel modes, ncu --set full of their traversal kernels (dram bytes next to
# the algorithmic bytes), packet-kernel baselines on config 2's frames.
set -u
TAG=${1:-r5d}
OUT=gpurun_out
mkdir -p $OUT
for C in 4 5; do
  for MODE in persistent phased; do
    RTBVH_TRACE_MODE=$MODE timeout 900 python bench.py --config $C --no-cpu --steps 10 --warmup 3 --e2e-steps 2 2> $OUT/${TAG}_c${C}_$MODE.err > $OUT/${TAG}_c${C}_$MODE.json
    python - <<PY
import json
try:
    d=json.load(open("$OUT/${TAG}_c${C}_$MODE.json")); print("config $C $MODE", round(d["value"],1), d["unit"], "e2e", round(d["e2e"]["value"],1))
except Exception as e: print("config $C $MODE FAILED", e)
PY
  done
done
{
for M in static persistent; do
  RTBVH_PACKET_MODE=$M timeout 300 python scripts/trace_ab.py --packets --name packet_$M 2>&1 | tail -1
  RTBVH_PACKET_MODE=$M timeout 300 python scripts/trace_ab.py --packets --any --name packet_$M 2>&1 | tail -1
  RTBVH_PACKET_MODE=$M timeout 300 python scripts/trace_ab.py --packets --bvh --name packet_$M 2>&1 | tail -1
done
RTBVH_TRACE_MODE=persistent timeout 300 python scripts/trace_ab.py --any --name single 2>&1 | tail -1
RTBVH_TRACE_MODE=persistent timeout 300 python scripts/trace_ab.py --bvh --name single 2>&1 | tail -1
RTBVH_TRACE_MODE=phased timeout 300 python scripts/trace_ab.py --bvh --name single 2>&1 | tail -1
} | tee $OUT/${TAG}_packets.txt
for C in 4 5; do
  echo "== ncu full config $C"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_single -s 2 -c 1 -f -o $OUT/${TAG}_prof_c$C \
      python bench.py --config $C --no-cpu --steps 2 --warmup 2 --e2e-steps 1 > $OUT/${TAG}_ncu_c$C.log 2>&1
  tail -1 $OUT/${TAG}_ncu_c$C.log | cut -c1-200
  ncu -i $OUT/${TAG}_prof_c$C.ncu-rep --page details > $OUT/${TAG}_trace_c${C}_details.txt 2>&1
done
