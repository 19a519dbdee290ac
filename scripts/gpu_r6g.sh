#!/bin/bash
set -u
TAG=${1:-r6g}
OUT=gpurun_out
mkdir -p $OUT
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29795 bench.py --gpus 4 --steps 60 --warmup 5 --no-cpu --e2e-steps 8 2> $OUT/${TAG}_n4.err > $OUT/${TAG}_n4.json
python - <<PY
import json
d=json.load(open("$OUT/${TAG}_n4.json"))
e=d["e2e"]
print("N=4", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "e2e", round(e["value"],1), "camera", round(e["camera_value"],1), d["config"].get("fused_gather_equals_all_gather"))
PY
