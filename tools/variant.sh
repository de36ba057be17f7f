#!/bin/bash
# usage: tools/variant.sh NAME "-DFLAG=1 ..."   -> build/libcrt_b200_NAME.so  (A/B builds, selected with CRT_B200_LIB)
set -e
cd "$(dirname "$0")/.."
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xcompiler -pthread -Iinclude -Icuda-raytracing-optimized_b200/csrc $2 \
  -shared cuda-raytracing-optimized_b200/csrc/renderer.cu cuda-raytracing-optimized_b200/csrc/wide_bvh.cpp -o build/libcrt_b200_$1.so
