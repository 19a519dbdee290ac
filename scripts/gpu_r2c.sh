#!/bin/bash
# Round-2 probe: per-kernel launch list of one 1 Mi-triangle binned-SAH build, pinned H2D/D2H bandwidth, e2e vs chunk size.
set -u
TAG=${1:-r2c}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python scripts/build_profile.py soup 2>&1 | tee $OUT/${TAG}_build1m.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/${TAG}_build_launches.csv python scripts/build_profile.py soup > $OUT/${TAG}_ncu_build.log 2>&1
tail -2 $OUT/${TAG}_ncu_build.log
python - <<'PY' 2>&1 | tee $OUT/${TAG}_pcie.txt
import torch, time
n = 256 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device="cuda")
h2 = torch.empty(n // 4, dtype=torch.uint8).pin_memory(); d2 = torch.empty(n // 4, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(both):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10):
        with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
        if both:
            with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / 10
run(False)
print("H2D alone GB/s", n / run(False) / 1e9)
print("H2D with concurrent D2H (1/4 size) GB/s", n / run(True) / 1e9)
PY
for c in 131072 262144 524288 1048576 2097152; do
  echo "chunk $c"; RTBVH_CHUNK_RAYS=$c timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --e2e-steps 20 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read());print(d['value'],d['e2e']['value'])"
done 2>&1 | tee $OUT/${TAG}_e2e_chunks.txt
