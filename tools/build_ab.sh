#!/bin/bash
# Build A/B variants of librx_b200.so: tools/build_ab.sh name1:"-DFLAG=1 ..." name2:"..."  -> build/ab/librx_<name>.so
set -e
cd "$(dirname "$0")/../openmmtools_b200/csrc"
out=../../build/ab; mkdir -p $out
ARCH="-gencode arch=compute_100a,code=sm_100a"
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  nvcc -O3 -std=c++17 -lineinfo $ARCH -Xcompiler -fPIC,-fvisibility=hidden --fmad=true $flags -c rx_mix.cu -o $out/rx_mix_$name.o
  nvcc $ARCH -shared -o $out/librx_$name.so rx_api.o $out/rx_mix_$name.o rx_dynamics.o -lcudart_static -ldl -lpthread -lrt
  echo built $out/librx_$name.so
done
