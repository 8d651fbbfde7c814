#!/bin/bash
# multi-GPU session on N GPUs of one box: tools/gpu_mgpu.sh N tag [check] [cfg5]
N=$1; TAG=$2; shift 2
mkdir -p gpurun_out
TR="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for a in "$@"; do
  if [ "$a" = check ]; then $TR --master-port 29511 tools/mgpu_check.py > gpurun_out/${TAG}_mgpu_check_${N}gpu.log 2>&1; grep "\[mgpu\]" gpurun_out/${TAG}_mgpu_check_${N}gpu.log; tail -n 3 gpurun_out/${TAG}_mgpu_check_${N}gpu.log | grep -i error; fi
done
$TR --master-port 29512 bench.py --gpus $N --steps 48 --warmup 8 --no-cpu-baseline > gpurun_out/${TAG}_cfg2_${N}gpu.json 2> gpurun_out/${TAG}_cfg2_${N}gpu.err
[ -n "$SKIP_CFG3" ] || $TR --master-port 29513 bench.py --gpus $N --config cfg3 --steps 24 --warmup 6 > gpurun_out/${TAG}_cfg3_${N}gpu.json 2> gpurun_out/${TAG}_cfg3_${N}gpu.err
$TR --master-port 29514 bench.py --gpus $N --config cfg4 --steps 4 --warmup 3 > gpurun_out/${TAG}_cfg4_${N}gpu.json 2> gpurun_out/${TAG}_cfg4_${N}gpu.err
for a in "$@"; do
  if [ "$a" = cfg5 ]; then $TR --master-port 29515 bench.py --gpus $N --config cfg5 --steps 1 --warmup 3 > gpurun_out/${TAG}_cfg5_${N}gpu.json 2> gpurun_out/${TAG}_cfg5_${N}gpu.err; fi
done
for f in gpurun_out/${TAG}_cfg*_${N}gpu.json; do python - <<PY
import json
try:
    d=json.load(open("$f")); print("$f", d["n_gpus"], "ms/step %.3f"%d["ms_per_step"], "value %.3f"%d["value"], d["unit"], "e2e %.3f"%d["e2e"]["value"])
except Exception as ex:
    print("$f", "unreadable:", ex)
PY
done
tail -n 3 gpurun_out/${TAG}_*_${N}gpu.err | grep -v "^$" | tail -20
