#!/bin/bash
set -u
TAG=${1:-r5l}
N=${2:-2}
OUT=gpurun_out
mkdir -p $OUT
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
  grep "step end times" $OUT/${TAG}_$NAME.err | cut -c1-400
}
run default RTBVH_BENCH_STEPTIMES=1
run nosampler RTBVH_BENCH_NOSAMPLER=1 RTBVH_BENCH_STEPTIMES=1
run nosampler_ring6 RTBVH_BENCH_NOSAMPLER=1 RTBVH_BENCH_RING=6 RTBVH_BENCH_STEPTIMES=1
