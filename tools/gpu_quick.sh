#!/bin/bash
# quick GPU iteration: parity tests + short bench.  usage: tools/gpu_quick.sh <tag> [pytest -k expr]
TAG=${1:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q ${2:+-k "$2"} > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
timeout 600 python bench.py --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?" >> gpurun_out/bench_$TAG.err
tail -25 gpurun_out/pytest_gpu_$TAG.log; tail -3 gpurun_out/bench_$TAG.err; cut -c1-400 gpurun_out/bench_$TAG.json
