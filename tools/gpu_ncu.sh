#!/bin/bash
# usage: bash tools/gpu_ncu.sh TAG KERNEL_REGEX [skip] [count] [extra bench args]
TAG=$1; K=$2; SKIP=${3:-8}; CNT=${4:-2}; shift 4
OUT=gpurun_out; mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$K" --launch-skip $SKIP -c $CNT \
   -f -o $OUT/${TAG}_ncu python bench.py --steps 2 --warmup 1 --min-seconds 0 --no-cpu-baseline --no-e2e --no-graph "$@" > $OUT/${TAG}_ncu_run.log 2>&1
ls -la $OUT/${TAG}_ncu.ncu-rep
