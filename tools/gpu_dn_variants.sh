#!/bin/bash
# denoiser timings of each tuning variant (VPT_LIB override): 1080p bench kernels + the 4K chain.  usage: tools/gpu_dn_variants.sh name1 name2 ...
mkdir -p gpurun_out
for v in base "$@"; do
  if [ "$v" = base ]; then unset VPT_LIB; else export VPT_LIB=$PWD/real-time-path-tracing-voxel-blocks_b200/libvpt_$v.so; fi
  python bench.py --steps 16 --warmup 4 --no-cpu-baseline > gpurun_out/dnvar_$v.json 2> gpurun_out/dnvar_$v.err
  python tools/bench_denoiser_4k.py 8 > gpurun_out/dnvar4k_$v.json 2>> gpurun_out/dnvar_$v.err
  python - <<PY
import json
d=json.load(open("gpurun_out/dnvar_$v.json")); e=json.load(open("gpurun_out/dnvar4k_$v.json"))
print("$v", "1080p chain %.4f"%d["roofline"]["denoiser_chain"]["ms"], [ (k["name"][:8],k["ms"]) for k in d["kernels"][2:8]], "| 4K %.4f"%e["denoise_total_ms"], [(p["name"][:8],p["ms"]) for p in e["passes"]])
PY
done
