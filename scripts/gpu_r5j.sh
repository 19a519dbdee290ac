#!/bin/bash
set -u
TAG=${1:-r5j}
N=${2:-2}
OUT=gpurun_out
mkdir -p $OUT
{
timeout 300 python scripts/trace_ab.py --name stream 2>&1 | tail -1
RTBVH_LIB=$PWD/rtbvh_b200/librtbvh_rs_nostream.so timeout 300 python scripts/trace_ab.py --name nostream 2>&1 | tail -1
} | tee $OUT/${TAG}_ab.txt
run() { # name, env...
  local NAME=$1; shift
  env "$@" timeout 900 python bench.py --gpus $N --steps 40 --warmup 5 --no-cpu --e2e-steps 2 2> $OUT/${TAG}_$NAME.err > $OUT/${TAG}_$NAME.json
  python - <<PY
import json
try:
    d=json.load(open("$OUT/${TAG}_$NAME.json"))
    print("N=$N $NAME", round(d["value"],1), "Mrays/s  ms/step", round(d["ms_per_step"],3), d["config"].get("fused_gather_equals_all_gather"))
except Exception as e:
    print("N=$N $NAME FAILED", e)
PY
}
run stream_push1 RTBVH_GATHER_PUSH=1
run stream_push0 RTBVH_GATHER_PUSH=0
run nostream_push1 RTBVH_GATHER_PUSH=1 RTBVH_LIB=$PWD/rtbvh_b200/librtbvh_rs_nostream.so
run stream_push1_ring4 RTBVH_GATHER_PUSH=1 RTBVH_BENCH_RING=4
