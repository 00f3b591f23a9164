#!/bin/bash
# tools/build_variant.sh NAME [-Dflags...]: tracks.cu compiled with extra flags and linked with the shipped objects into
# variants/libssrs_NAME.so (kernel A/B experiments: SSRS_B200_LIB=variants/libssrs_NAME.so python bench.py ...)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --extended-lambda -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -Xptxas -v -cudart shared -fmad=false "$@" -c ssrs_b200/csrc/tracks.cu -o variants/tracks_$name.o 2> variants/tracks_$name.log || { tail -5 variants/tracks_$name.log; exit 1; }
objs=$(ls ssrs_b200/build/*.o | grep -v tracks.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart shared -o variants/libssrs_$name.so variants/tracks_$name.o $objs -Xlinker -rpath -Xlinker /usr/local/cuda/lib64 -ldl
