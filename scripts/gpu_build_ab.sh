#!/bin/bash
# Builder parity tests + traversal A/B (LDG.256 vs LDG.128) + per-kernel launch list of one 1 Mi-triangle build.
set -u
TAG=${1:-r1c}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -30 | tee $OUT/${TAG}_pytest_gpu.txt
echo "== bench ld256"
timeout 600 python bench.py --steps 40 --warmup 5 2> $OUT/${TAG}_bench_ld256.err | tee $OUT/${TAG}_bench_ld256.json | cut -c1-1500
tail -3 $OUT/${TAG}_bench_ld256.err
echo "== bench ld128"
RTBVH_LIB=$PWD/rtbvh_b200/librtbvh_rs_ld128.so timeout 600 python bench.py --steps 40 --warmup 5 2> $OUT/${TAG}_bench_ld128.err | tee $OUT/${TAG}_bench_ld128.json | cut -c1-400
echo "== ncu launch list of a 1Mi build"
cat > /tmp/build1m.py <<'PY'
import sys; sys.path.insert(0, '.')
from rtbvh_b200 import api, workloads as W
tris = W.soup(1 << 20)
for kind in (api.BINNED_SAH, api.LOCALLY_ORDERED_CLUSTERED):
    b = api.build_triangles(tris, kind, 1); print(kind, api.last_build_stats(), b.rt.node_count)
    m = api.Mbvh.construct(b); print('collapse', api.last_build_stats(), m.rt.node_count)
PY
timeout 600 python /tmp/build1m.py 2>&1 | tee $OUT/${TAG}_build1m.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/${TAG}_build_launches.csv python /tmp/build1m.py > $OUT/${TAG}_ncu_build.log 2>&1
tail -2 $OUT/${TAG}_ncu_build.log
