#!/bin/bash
set -u
TAG=${1:-r6b}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29781 bench.py --gpus 8 --steps 60 --warmup 5 --no-cpu --e2e-steps 8 2> $OUT/${TAG}_n8.err > $OUT/${TAG}_n8.json
python - <<PY
import json
d=json.load(open("$OUT/${TAG}_n8.json"))
e=d["e2e"]
print("N=8", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "e2e", round(e["value"],1), "camera", round(e["camera_value"],1), d["config"].get("fused_gather_equals_all_gather"), d["config"].get("replication"))
PY
tail -2 $OUT/${TAG}_n8.err | cut -c1-200
