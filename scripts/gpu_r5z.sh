#!/bin/bash
set -u
TAG=${1:-r5z}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python bench.py --steps 40 --warmup 5 2> $OUT/${TAG}_n1.err > $OUT/${TAG}_n1.json
python - <<PY
import json
d=json.load(open("$OUT/${TAG}_n1.json"))
print("N=1", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "packet4", d["config"].get("packet4"), "cpu", d["cpu_baseline"] and round(d["cpu_baseline"]["value"],2), d["cpu_baseline"] and d["cpu_baseline"].get("binned_sah_build_ms_per_mtri"), "frac", round(d["roofline"]["frac"],3), d["roofline"]["traffic"])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29777 bench.py --gpus 2 --steps 40 --warmup 5 2> $OUT/${TAG}_n2.err > $OUT/${TAG}_n2.json
python - <<PY
import json
d=json.load(open("$OUT/${TAG}_n2.json"))
print("N=2", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "camera", round(d["e2e"]["camera_value"],1), "packet4", d["config"].get("packet4",{}).get("value"), "cpu", d["cpu_baseline"], "frac", d["roofline"] and round(d["roofline"]["frac"],3), d["config"].get("replication"), d["config"].get("fused_gather_equals_all_gather"))
PY
tail -2 $OUT/${TAG}_n2.err | cut -c1-300
