#!/bin/bash
# One GPU-box session: parity tests, smoke, bench, ncu launch list + one full capture of the top kernel.
# Usage (from the repo root, under gpurun):  bash scripts/gpu_round.sh <tag> [bench steps]
set -u
TAG=${1:-r1}
STEPS=${2:-40}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/${TAG}_gpu.txt 2>&1
nproc >> $OUT/${TAG}_gpu.txt
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 | tee $OUT/${TAG}_pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/${TAG}_smoke.txt
echo "== bench"; timeout 900 python bench.py --steps $STEPS --warmup 5 2> $OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench.json
tail -5 $OUT/${TAG}_bench.err
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"trace_|camera_rays|gather_tris|ray_keys|peer_barrier" -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu --e2e-steps 1 > $OUT/${TAG}_ncu_launches.log 2>&1
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_single -s 3 -c 1 -f -o $OUT/${TAG}_prof \
    python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > $OUT/${TAG}_ncu_full.log 2>&1
ls -la $OUT | tail -20
