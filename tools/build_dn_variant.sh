#!/bin/bash
# A/B builds of the denoiser translation units only: tools/build_dn_variant.sh <name> "<extra nvcc -D flags>"
#  -> real-time-path-tracing-voxel-blocks_b200/libvpt_<name>.so (the other objects come from build/; run `make` first)
set -e
NAME=$1; EXTRA=$2
cd "$(dirname "$0")/../real-time-path-tracing-voxel-blocks_b200"
B=build_$NAME; mkdir -p $B
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -ccbin /usr/bin/g++ -I../include"
for f in denoise dn_tiles; do $NV -prec-div=false -prec-sqrt=false -ftz=true $EXTRA -c csrc/vpt_$f.cu -o $B/vpt_$f.o & done
$NV -prec-div=false -prec-sqrt=false $EXTRA -c csrc/vpt_temporal.cu -o $B/vpt_temporal.o &
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -ccbin /usr/bin/g++ -o libvpt_$NAME.so build/vpt_wave.o build/vpt_dda.o $B/vpt_denoise.o $B/vpt_dn_tiles.o $B/vpt_temporal.o build/vpt_sky.o build/vpt_grid.o build/vpt_api.o build/vpt_host.o build/vpt_lights.o -ldl
echo built libvpt_$NAME.so
