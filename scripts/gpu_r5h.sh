#!/bin/bash
# Round 5, session h (N GPUs): fused-gather + replication tests on distinct GPUs, bench at N GPUs with the chunk-wise push on / off.
set -u
TAG=${1:-r5h}
N=${2:-2}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi topo -m > $OUT/${TAG}_topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_fused_gather.py tests/test_gpu_replicate.py -x -q -m gpu 2>&1 | tail -6 | tee $OUT/${TAG}_pytest.txt
for PUSH in 1 0; do
  RTBVH_GATHER_PUSH=$PUSH timeout 900 python bench.py --gpus $N --steps 40 --warmup 5 --no-cpu --e2e-steps 10 2> $OUT/${TAG}_bench_n${N}_push$PUSH.err > $OUT/${TAG}_bench_n${N}_push$PUSH.json
  python - <<PY
import json
try:
    d=json.load(open("$OUT/${TAG}_bench_n${N}_push$PUSH.json"))
    print("N=$N push=$PUSH", round(d["value"],1), "Mrays/s  e2e", round(d["e2e"]["value"],1), "camera", round(d["e2e"]["camera_value"],1), d["config"].get("nvlink_gpu0"), d["config"].get("replication"), d["config"].get("fused_gather_equals_all_gather"))
except Exception as e:
    print("N=$N push=$PUSH FAILED", e)
PY
  tail -3 $OUT/${TAG}_bench_n${N}_push$PUSH.err
done
timeout 600 python bench.py --gpus 1 --steps 40 --warmup 5 --no-cpu --e2e-steps 10 2> $OUT/${TAG}_bench_n1.err > $OUT/${TAG}_bench_n1.json
python -c "
import json; d=json.load(open('$OUT/${TAG}_bench_n1.json')); print('N=1', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'camera', round(d['e2e']['camera_value'],1))"
