#!/bin/bash
set -u
TAG=${1:-r6d}
OUT=gpurun_out
mkdir -p $OUT
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29791 bench.py --config 4 --gpus 8 --steps 10 --warmup 3 --no-cpu --e2e-steps 3 2> $OUT/${TAG}_c4_n8.err > $OUT/${TAG}_c4_n8.json
python - <<PY
import json
d=json.load(open("$OUT/${TAG}_c4_n8.json"))
print("config 4 N=8", round(d["value"],1), d["unit"], "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1))
PY
tail -2 $OUT/${TAG}_c4_n8.err | cut -c1-200
