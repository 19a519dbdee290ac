#!/bin/bash
set -u
TAG=${1:-r6a}
N=${2:-2}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_fused_gather.py -x -q -m gpu 2>&1 | tail -4
for I in 1 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2977$I bench.py --gpus $N --steps 60 --warmup 5 --no-cpu --e2e-steps 4 2> $OUT/${TAG}_n$N.err > $OUT/${TAG}_n$N.json
python - <<PY
import json
d=json.load(open("$OUT/${TAG}_n$N.json"))
print("N=$N", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), d["config"].get("fused_gather_equals_all_gather"))
PY
done
timeout 600 python bench.py --steps 60 --warmup 5 --no-cpu --e2e-steps 4 2> $OUT/${TAG}_n1.err > $OUT/${TAG}_n1.json
python -c "
import json; d=json.load(open('$OUT/${TAG}_n1.json')); print('N=1', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1))"
