#!/bin/bash
# Round 5, session n (8 GPUs): PCIe ceiling at 1/2/4/8 ranks, fused gather isolation at 8 ranks, bench at 8 GPUs with the
# chunk-wise push on / off, bench at 4 GPUs.
set -u
TAG=${1:-r5n}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi topo -m > $OUT/${TAG}_topo.txt 2>&1
nproc > $OUT/${TAG}_nproc.txt; lscpu | grep -E "NUMA|Socket|Model name" >> $OUT/${TAG}_nproc.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for N in 1 2 4 8; do
  timeout 200 $TR --nproc-per-node $N --master-port $((29600+N)) scripts/pcie_ceiling.py 2>/dev/null | grep "^{" | tee -a $OUT/${TAG}_pcie_ceiling.jsonl | cut -c1-420
done
for P in 1 0; do RTBVH_GATHER_PUSH=$P STEPS=30 timeout 300 $TR --nproc-per-node 8 --master-port 29633 scripts/gather_ab.py 2>&1 | grep "^N="; done | tee $OUT/${TAG}_gather_ab.txt
run() { # name, n, env...
  local NAME=$1; shift; local N=$1; shift
  env "$@" timeout 600 python bench.py --gpus $N --steps 40 --warmup 5 --no-cpu --e2e-steps 10 2> $OUT/${TAG}_$NAME.err > $OUT/${TAG}_$NAME.json
  python - <<PY
import json
try:
    d=json.load(open("$OUT/${TAG}_$NAME.json"))
    e=d["e2e"]
    print("$NAME", round(d["value"],1), "Mrays/s  ms/step", round(d["ms_per_step"],3), "e2e od", round(e["value"],1), "rtray", round(e["rtray_async_value"],1), "camera", round(e["camera_value"],1), d["config"].get("fused_gather_equals_all_gather"), d["config"].get("host_numa"))
except Exception as e:
    print("$NAME FAILED", e)
PY
}
run n8_push1 8 RTBVH_GATHER_PUSH=1
run n8_push0 8 RTBVH_GATHER_PUSH=0
run n4_push1 4 RTBVH_GATHER_PUSH=1
run n1 1
