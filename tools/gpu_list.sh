#!/bin/bash
# bench + ncu launch list of the same command.  usage: tools/gpu_list.sh <tag>
TAG=${1:-l}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 4 --no-cpu-baseline"
$CMD > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
cut -c1-300 gpurun_out/bench_$TAG.json; tail -2 gpurun_out/bench_$TAG.err
