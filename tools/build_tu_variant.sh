#!/bin/bash
# A/B builds of ONE translation unit: tools/build_tu_variant.sh <name> <wave|dda|denoise|dn_tiles|temporal> "<-D flags>"
#   -> real-time-path-tracing-voxel-blocks_b200/libvpt_<name>.so (all other objects from build/; flags per TU as in the Makefile)
set -e
NAME=$1; TU=$2; EXTRA=$3
cd "$(dirname "$0")/../real-time-path-tracing-voxel-blocks_b200"
B=build_$NAME; mkdir -p $B
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -ccbin /usr/bin/g++ -I../include -prec-div=false -prec-sqrt=false"
case $TU in
  wave) FL="-use_fast_math";;
  denoise|dn_tiles) FL="-ftz=true";;
  *) FL="";;
esac
$NV $FL $EXTRA -c csrc/vpt_$TU.cu -o $B/vpt_$TU.o
OBJ=$(ls build/*.o | grep -v vpt_$TU.o)
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -ccbin /usr/bin/g++ -o libvpt_$NAME.so $OBJ $B/vpt_$TU.o -ldl
echo built libvpt_$NAME.so
