#!/bin/bash
# Round 4, session c: pooled pinned host mirrors + in-place collapse (drop-in build path), software-pipelined node fetch,
# 8x8 tile work order.
set -u
TAG=${1:-r4c}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu (default lib)"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee $OUT/${TAG}_pytest_gpu.txt
echo "== pytest traversal (pipe)"
RTBVH_LIB=$PWD/rtbvh_b200/librtbvh_rs_pipe.so timeout 1500 python -m pytest tests/test_gpu_traversal.py tests/test_zz_gpu_golden.py -x -q -m gpu 2>&1 | tail -3 | tee $OUT/${TAG}_pytest_gpu_pipe.txt
bench() {  # name, env...
  local NAME=$1; shift
  env "$@" timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu --e2e-steps 4 2> $OUT/${TAG}_$NAME.err > $OUT/${TAG}_$NAME.json
  python - <<PY
import json
try:
    d=json.load(open("$OUT/${TAG}_$NAME.json"))
    b=d["config"]["build"]
    print("$NAME", round(d["value"],1), "Mrays/s  e2e", round(d["e2e"]["value"],1), d["e2e"]["host_equals_resident"], "| build dev", round(b["binned_sah_ms_per_mtri"],2), "incl", round(b["binned_sah_ms_per_mtri_incl_h2d_d2h"],2), "collapse dev", round(b["collapse_device_ms"],2), "incl", round(b["collapse_ms_incl_h2d_d2h"],2))
except Exception as e:
    print("$NAME", "FAILED", e)
PY
}
bench persistent_notile RTBVH_TRACE_MODE=persistent RTBVH_BENCH_TILING=0
bench persistent_tile RTBVH_TRACE_MODE=persistent RTBVH_BENCH_TILING=1
bench phased_notile RTBVH_BENCH_TILING=0
bench phased_tile RTBVH_BENCH_TILING=1
for LIB in rtbvh_b200/librtbvh_rs_*.so; do
  bench $(basename $LIB .so)_tile RTBVH_LIB=$PWD/$LIB RTBVH_BENCH_TILING=1
done
bench pipe_notile RTBVH_LIB=$PWD/rtbvh_b200/librtbvh_rs_pipe.so RTBVH_BENCH_TILING=0
for V in pipe t21; do
echo "== ncu full ($V, tiling)"
RTBVH_LIB=$PWD/rtbvh_b200/librtbvh_rs_$V.so timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_single -s 3 -c 1 -f -o $OUT/${TAG}_prof_$V \
    python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > $OUT/${TAG}_ncu_$V.log 2>&1
tail -1 $OUT/${TAG}_ncu_$V.log | cut -c1-200
done
