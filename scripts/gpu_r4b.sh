#!/bin/bash
# Round 4, session b: deferred stores / cheaper leaf handling / eager tri policies / staged top of the tree (RTB_TOPK).
set -u
TAG=${1:-r4b}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu (default lib)"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee $OUT/${TAG}_pytest_gpu.txt
echo "== pytest traversal + dynamic (top341b896)"
RTBVH_LIB=$PWD/rtbvh_b200/librtbvh_rs_top341b896.so timeout 1500 python -m pytest tests/test_gpu_traversal.py tests/test_gpu_dynamic.py tests/test_zz_gpu_golden.py -x -q -m gpu 2>&1 | tail -5 | tee $OUT/${TAG}_pytest_gpu_top.txt
bench() {  # name, env...
  local NAME=$1; shift
  env "$@" timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu --e2e-steps 4 2> $OUT/${TAG}_$NAME.err > $OUT/${TAG}_$NAME.json
  python - <<PY
import json
try:
    d=json.load(open("$OUT/${TAG}_$NAME.json"))
    print("$NAME", round(d["value"],1), "Mrays/s  e2e", round(d["e2e"]["value"],1), d["e2e"]["host_equals_resident"])
except Exception as e:
    print("$NAME", "FAILED", e)
PY
}
bench persistent RTBVH_TRACE_MODE=persistent
bench phased RTBVH_TRACE_MODE=phased
for LIB in rtbvh_b200/librtbvh_rs_*.so; do
  bench $(basename $LIB .so) RTBVH_LIB=$PWD/$LIB
done
for V in top341b896 e32; do
echo "== ncu full ($V)"
RTBVH_LIB=$PWD/rtbvh_b200/librtbvh_rs_$V.so timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_single -s 3 -c 1 -f -o $OUT/${TAG}_prof_$V \
    python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > $OUT/${TAG}_ncu_$V.log 2>&1
tail -1 $OUT/${TAG}_ncu_$V.log | cut -c1-200
done
