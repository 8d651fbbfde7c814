#!/bin/bash
# A/B builds of the DDA engine only: tools/build_dda_variant.sh <name> "<-D flags>" -> libvpt_<name>.so (other objects from build/)
set -e
NAME=$1; EXTRA=$2
cd "$(dirname "$0")/../real-time-path-tracing-voxel-blocks_b200"
B=build_$NAME; mkdir -p $B
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -ccbin /usr/bin/g++ -I../include \
  -prec-div=false -prec-sqrt=false $EXTRA -c csrc/vpt_dda.cu -o $B/vpt_dda.o
OBJ=$(ls build/*.o | grep -v vpt_dda.o)
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -ccbin /usr/bin/g++ -o libvpt_$NAME.so $OBJ $B/vpt_dda.o -ldl
echo built libvpt_$NAME.so
