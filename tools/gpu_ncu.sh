#!/bin/bash
# ncu --set full capture of selected kernels of one steady-state frame.  usage: tools/gpu_ncu.sh <tag> <kernel regex> <skip> <count>
TAG=${1:-n}; RE=${2:-dda}; SKIP=${3:-0}; CNT=${4:-12}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 4 --no-cpu-baseline"
$CMD > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"$RE" -s $SKIP -c $CNT -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
tail -3 gpurun_out/ncu_full_$TAG.log | cut -c1-200
