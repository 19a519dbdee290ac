#!/bin/bash
# Round 5, session a (re-entry): full GPU suite, smoke under ncu, default bench, ncu full of the default traversal kernel,
# launch list of the bench, then the other configs.
set -u
TAG=${1:-r5a}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/${TAG}_gpu.txt 2>&1
nproc >> $OUT/${TAG}_gpu.txt
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu --durations=12 2>&1 | tail -30 | tee $OUT/${TAG}_pytest_gpu.txt
echo "== smoke under ncu (launch list)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 50 --csv --log-file $OUT/${TAG}_smoke_launches.csv \
    python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke_ncu.log 2>&1
echo "rc=$?"; tail -2 $OUT/${TAG}_smoke_ncu.log; grep -c trace_ $OUT/${TAG}_smoke_launches.csv
echo "== bench (default = config 2)"
timeout 900 python bench.py --steps 40 --warmup 5 2> $OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench.json | cut -c1-1500
tail -3 $OUT/${TAG}_bench.err
echo "== ncu full (default trace kernel)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_single -s 3 -c 1 -f -o $OUT/${TAG}_prof \
    python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > $OUT/${TAG}_ncu_full.log 2>&1
tail -2 $OUT/${TAG}_ncu_full.log | cut -c1-200
ncu -i $OUT/${TAG}_prof.ncu-rep --page details > $OUT/${TAG}_trace_details.txt 2>&1
echo "== launch list of the bench"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --e2e-steps 1 > $OUT/${TAG}_launches.log 2>&1
echo "rc=$?"
for C in 1 3 4 5; do
  echo "== bench --config $C"
  timeout 1200 python bench.py --config $C 2> $OUT/${TAG}_bench_c$C.err | tee $OUT/${TAG}_bench_c$C.json | cut -c1-700
  tail -3 $OUT/${TAG}_bench_c$C.err
done
