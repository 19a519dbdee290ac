#!/bin/bash
# Round 5, final session: full GPU suite, smoke under ncu, default bench (both arms), ncu --set full of the default single-ray
# kernel and of the packet kernel, launch list, the other configs (both arms for config 1).
set -u
TAG=${1:-r5s}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/${TAG}_gpu.txt 2>&1
nproc >> $OUT/${TAG}_gpu.txt
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee $OUT/${TAG}_pytest_gpu.txt
echo "== smoke under ncu (launch list)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 50 --csv --log-file $OUT/${TAG}_smoke_launches.csv \
    python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke_ncu.log 2>&1
echo "rc=$?"; tail -1 $OUT/${TAG}_smoke_ncu.log; grep -c trace_ $OUT/${TAG}_smoke_launches.csv
echo "== bench (default = config 2)"
timeout 900 python bench.py 2> $OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench.json | cut -c1-500
tail -2 $OUT/${TAG}_bench.err
echo "== bench --impl reference"
timeout 900 python bench.py --impl reference --steps 10 --warmup 3 2> $OUT/${TAG}_bench_ref.err | tee $OUT/${TAG}_bench_ref.json | cut -c1-400
echo "== ncu full (default single-ray kernel)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_single -s 3 -c 1 -f -o $OUT/${TAG}_prof \
    python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > $OUT/${TAG}_ncu_full.log 2>&1
ncu -i $OUT/${TAG}_prof.ncu-rep --page details > $OUT/${TAG}_trace_details.txt 2>&1
grep -E "Duration|Executed Ipc Active|Avg. Active Threads|L2 Hit Rate|DRAM Throughput" $OUT/${TAG}_trace_details.txt | head -6
echo "== ncu full (packet kernel)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:packet_lane -s 3 -c 1 -f -o $OUT/${TAG}_prof_packet \
    python scripts/trace_ab.py --packets --steps 2 --warmup 3 > $OUT/${TAG}_ncu_packet.log 2>&1
ncu -i $OUT/${TAG}_prof_packet.ncu-rep --page details > $OUT/${TAG}_packet_details.txt 2>&1
grep -E "Duration|Executed Ipc Active|Avg. Active Threads|L2 Hit Rate|DRAM Throughput|Registers Per" $OUT/${TAG}_packet_details.txt | head -6
echo "== launch list of the bench"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu --e2e-steps 1 > $OUT/${TAG}_launches.log 2>&1
echo "rc=$?"
for C in 1 3 4 5; do
  echo "== bench --config $C"
  timeout 1200 python bench.py --config $C 2> $OUT/${TAG}_bench_c$C.err | tee $OUT/${TAG}_bench_c$C.json | cut -c1-300
  tail -2 $OUT/${TAG}_bench_c$C.err
done
echo "== bench --config 1 --impl reference"
timeout 600 python bench.py --config 1 --impl reference --steps 3 --warmup 1 2> $OUT/${TAG}_ref_c1.err | tee $OUT/${TAG}_ref_c1.json | cut -c1-300
