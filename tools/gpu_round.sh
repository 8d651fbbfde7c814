#!/bin/bash
# One gpurun call: GPU parity tests, bench (both arms), ncu launch list + full capture of the hot kernels.
# usage: tools/gpu_round.sh <tag>   (outputs under gpurun_out/)
TAG=${1:-r1}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_$TAG.log
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?" >> gpurun_out/bench_$TAG.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
CMD="python bench.py --steps 2 --warmup 6 --no-cpu-baseline"
$CMD > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
$CMD > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"genKernel|dda|shade|accumulate|prepKernel|temporalKernel|historyFix|historyClamp|atrous|firefly" -s 115 -c 23 -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
tail -3 gpurun_out/pytest_gpu_$TAG.log; tail -2 gpurun_out/smoke_$TAG.log; cat gpurun_out/bench_$TAG.json | cut -c1-600
