#!/bin/bash
# bench each tuning variant (VPT_LIB override).  usage: tools/gpu_variants.sh name1 name2 ...
mkdir -p gpurun_out
for v in base "$@"; do
  if [ "$v" = base ]; then unset VPT_LIB; else export VPT_LIB=$PWD/real-time-path-tracing-voxel-blocks_b200/libvpt_$v.so; fi
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_var_$v.json 2> gpurun_out/bench_var_$v.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_var_$v.json"))
print("$v", "ms/frame %.3f"%d["ms_per_step"], "trace", d["trace"]["ms"], "dda", d["trace"]["dda_ms"], "shade", d["trace"]["shade_ms"], [ (k["name"],k["ms"]) for k in d["kernels"][2:]])
PY
done
