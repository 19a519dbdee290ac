#!/bin/bash
# Round 5, session i (N GPUs): fused gather tests, bench at N GPUs: chunk-wise push on / off x tiling on / off.
set -u
TAG=${1:-r5i}
N=${2:-2}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_fused_gather.py -x -q -m gpu 2>&1 | tail -6 | tee $OUT/${TAG}_pytest.txt
for PUSH in 1 0; do
 for TIL in 1 0; do
  RTBVH_BENCH_TILING=$TIL RTBVH_GATHER_PUSH=$PUSH timeout 900 python bench.py --gpus $N --steps 40 --warmup 5 --no-cpu --e2e-steps 4 2> $OUT/${TAG}_bench_n${N}_push${PUSH}_til$TIL.err > $OUT/${TAG}_bench_n${N}_push${PUSH}_til$TIL.json
  python - <<PY
import json
try:
    d=json.load(open("$OUT/${TAG}_bench_n${N}_push${PUSH}_til$TIL.json"))
    print("N=$N push=$PUSH tiling=$TIL", round(d["value"],1), "Mrays/s  ms/step", round(d["ms_per_step"],3), "camera e2e", round(d["e2e"]["camera_value"],1), d["config"].get("fused_gather_equals_all_gather"))
except Exception as e:
    print("N=$N push=$PUSH FAILED", e)
PY
 done
done
RTBVH_BENCH_TILING=1 timeout 900 python bench.py --gpus $N --steps 40 --warmup 5 --no-cpu --e2e-steps 4 --gather nccl 2> $OUT/${TAG}_bench_n${N}_nccl.err > $OUT/${TAG}_bench_n${N}_nccl.json
python -c "
import json; d=json.load(open('$OUT/${TAG}_bench_n${N}_nccl.json')); print('N=$N nccl gather', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3))"
RTBVH_BENCH_TILING=1 timeout 900 python bench.py --gpus $N --steps 40 --warmup 5 --no-cpu --e2e-steps 4 --no-gather 2> $OUT/${TAG}_bench_n${N}_nogather.err > $OUT/${TAG}_bench_n${N}_nogather.json
python -c "
import json; d=json.load(open('$OUT/${TAG}_bench_n${N}_nogather.json')); print('N=$N no gather', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3))"
nvidia-smi nvlink -gt d -i 0 2>&1 | head -8
