#!/bin/bash
set -u
TAG=${1:-r5x}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python scripts/sanitize_smoke.py 2>&1 | tail -3
for TOOL in memcheck racecheck; do
  echo "== compute-sanitizer --tool $TOOL"
  timeout 1500 compute-sanitizer --tool $TOOL --log-file $OUT/${TAG}_$TOOL.log python scripts/sanitize_smoke.py 2>&1 | tail -2
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error|Race" $OUT/${TAG}_$TOOL.log | sort | uniq -c | head -12
done
