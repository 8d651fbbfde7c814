#!/bin/bash
# A/B tuning builds: tools/build_variant.sh <name> "<extra nvcc -D flags>"  -> real-time-path-tracing-voxel-blocks_b200/libvpt_<name>.so
set -e
NAME=$1; EXTRA=$2
cd "$(dirname "$0")/../real-time-path-tracing-voxel-blocks_b200"
B=build_$NAME; mkdir -p $B
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -ccbin /usr/bin/g++ -I../include"
$NV -use_fast_math $EXTRA -c csrc/vpt_wave.cu -o $B/vpt_wave.o &
for f in denoise; do $NV -prec-div=false -prec-sqrt=false -ftz=true $EXTRA -c csrc/vpt_$f.cu -o $B/vpt_$f.o & done
for f in dda api; do $NV -prec-div=false -prec-sqrt=false $EXTRA -c csrc/vpt_$f.cu -o $B/vpt_$f.o & done
$NV -prec-div=false -prec-sqrt=false $EXTRA -c csrc/vpt_temporal.cu -o $B/vpt_temporal.o &
$NV -fmad=false $EXTRA -c csrc/vpt_grid.cu -o $B/vpt_grid.o &
$NV -fmad=false -DVPT_FAST_MATH=0 $EXTRA -c csrc/vpt_sky.cu -o $B/vpt_sky.o &
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -ccbin /usr/bin/g++ -o libvpt_$NAME.so $B/vpt_wave.o $B/vpt_dda.o $B/vpt_denoise.o $B/vpt_temporal.o $B/vpt_sky.o $B/vpt_grid.o $B/vpt_api.o build/vpt_host.o build/vpt_lights.o -ldl
echo built libvpt_$NAME.so
