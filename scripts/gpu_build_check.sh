#!/bin/bash
# Builder iteration: parity tests of the builders, bench's build timing, per-kernel launch list of one 1 Mi build.
set -u
TAG=${1:-bchk}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_build.py tests/test_gpu_dynamic.py -x -q -m gpu 2>&1 | tail -8
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --e2e-steps 2 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read());print(d['config']['build'])" | tee $OUT/${TAG}_build.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/${TAG}_build_launches.csv python scripts/build_profile.py soup > $OUT/${TAG}_ncu_build.log 2>&1
tail -1 $OUT/${TAG}_ncu_build.log
