#!/bin/bash
# Round 4, session a: parity (phased kernel is the default now; blocking-launch test), smoke under ncu, A/B of the trace
# modes and of the phase-policy variants, ncu full capture of the phased kernel.
set -u
TAG=${1:-r4a}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/${TAG}_gpu.txt 2>&1
nproc >> $OUT/${TAG}_gpu.txt
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 | tee $OUT/${TAG}_pytest_gpu.txt
echo "== smoke under ncu (launch list)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 50 --csv --log-file $OUT/${TAG}_smoke_launches.csv \
    python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke_ncu.log 2>&1
echo "rc=$?"; tail -2 $OUT/${TAG}_smoke_ncu.log; grep -c trace_ $OUT/${TAG}_smoke_launches.csv
for MODE in persistent phased; do
  RTBVH_TRACE_MODE=$MODE timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu --e2e-steps 10 2> $OUT/${TAG}_bench_$MODE.err > $OUT/${TAG}_bench_$MODE.json
  python - <<PY
import json
d=json.load(open("$OUT/${TAG}_bench_$MODE.json"))
print("$MODE", round(d["value"],1), "Mrays/s  e2e", round(d["e2e"]["value"],1), "sync", round(d["e2e"]["sync_call_value"],1), d["e2e"]["host_equals_resident"])
PY
done
for LIB in rtbvh_b200/librtbvh_rs_*.so; do
  NAME=$(basename $LIB .so)
  RTBVH_LIB=$PWD/$LIB timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu --e2e-steps 2 2> $OUT/${TAG}_$NAME.err > $OUT/${TAG}_$NAME.json
  python - <<PY
import json
d=json.load(open("$OUT/${TAG}_$NAME.json"))
print("$NAME", round(d["value"],1), "Mrays/s  e2e", round(d["e2e"]["value"],1))
PY
done
echo "== ncu full (phased)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_single -s 3 -c 1 -f -o $OUT/${TAG}_prof \
    python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > $OUT/${TAG}_ncu_full.log 2>&1
tail -2 $OUT/${TAG}_ncu_full.log | cut -c1-200
echo "== full bench (default)"
timeout 900 python bench.py --steps 40 --warmup 5 2> $OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench.json | cut -c1-600
