#!/bin/bash
# A/B of library variants on the benchmark frame, same box, back to back (run on the GPU box through gpurun).
# usage: tools/ab.sh NAME [NAME ...]   -- NAME as given to tools/variant.sh (build/libcrt_b200_NAME.so); "base" = build/libcrt_b200.so
cd "$(dirname "$0")/.."
for v in "$@" base; do
  lib=build/libcrt_b200_$v.so
  [ "$v" = base ] && lib=build/libcrt_b200.so
  echo "variant: $v"
  CRT_B200_LIB=$PWD/$lib python tools/render_once.py --ns 100 --steps 3 2>/dev/null | tail -2 | cut -c1-70
done
