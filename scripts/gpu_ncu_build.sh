#!/bin/bash
# ncu --set full of the binned-SAH builder kernels on ONE 1 Mi-triangle build (scripts/build_once.py):
#   A: the small-subtree kernel (largest single item of a build), with source;
#   B: the level-loop kernels (span bin, warp tasks, split, emit, the partition scan), first 48 instances, no source.
# Text summaries are written next to the reports so that they can be read without ncu.
set -u
TAG=${1:-ncu_build}
OUT=gpurun_out
mkdir -p $OUT
timeout 150 ncu --set full --clock-control none --import-source on -k regex:sah_small_kernel -c 1 -f -o $OUT/${TAG}_small \
    python scripts/build_once.py > $OUT/${TAG}_small.log 2>&1
timeout 60 ncu -i $OUT/${TAG}_small.ncu-rep --page details > $OUT/${TAG}_small_details.txt 2>&1
timeout 200 ncu --set full --clock-control none -k 'regex:sah_bin_kernel|sah_warp_task_kernel|sah_split_kernel|sah_emit_kernel|ScanByKey' \
    -c 48 -f -o $OUT/${TAG}_levels python scripts/build_once.py > $OUT/${TAG}_levels.log 2>&1
timeout 60 ncu -i $OUT/${TAG}_levels.ncu-rep --page raw --csv --metrics \
gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,\
gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,\
lts__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed,\
sm__warps_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,\
launch__grid_size,launch__block_size,launch__registers_per_thread,smsp__issue_active.avg.pct_of_peak_sustained_active \
    > $OUT/${TAG}_levels_raw.csv 2>&1
tail -2 $OUT/${TAG}_small.log $OUT/${TAG}_levels.log | cut -c1-200
ls -la $OUT | grep ${TAG}
