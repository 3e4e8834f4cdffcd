#!/bin/bash
# build_variant.sh NAME "-DE3_G=4 ..." -> build/libporrt_NAME.so (A/B builds of the kernels; select with PORRT_B200_LIB)
set -e
cd "$(dirname "$0")/../po_rrt_b200/csrc"
mkdir -p ../../build/$1
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC,-fvisibility=hidden -ccbin /usr/bin/g++"
for f in ctx map edge3 edge4 nn nn_tile graph refine; do /usr/local/cuda/bin/nvcc $FLAGS $2 -c $f.cu -o ../../build/$1/$f.o & done; wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../build/libporrt_$1.so ../../build/$1/*.o -lcudart_static -lpthread -ldl -lrt
echo built build/libporrt_$1.so
