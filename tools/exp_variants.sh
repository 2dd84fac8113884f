#!/bin/bash
# Developer tool: builds timing-experiment variants of the 20-state walk kernel (plg_walk_aa.cu) (parts of the
# operation switched off by AW_EXP_* macros; results are wrong on purpose) into tools/exp/.
set -e
cd "$(dirname "$0")/../libpll_b200/csrc"
mkdir -p ../../tools/exp
OBJS=$(ls build/*.o | grep -v plg_walk_aa)
for v in "$@"; do
  flags=$(echo "$v" | tr '+' '\n' | sed 's/^/-DAW_EXP_/' | tr '\n' ' ')
  /usr/local/cuda/bin/nvcc $flags -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false \
    -Xcompiler -fPIC,-fvisibility=hidden -I../../include -Igpu -Ihost -c gpu/plg_walk_aa.cu -o ../../tools/exp/aa_$v.o
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o ../../tools/exp/lib_$v.so \
    $OBJS ../../tools/exp/aa_$v.o -Xlinker -Bsymbolic -lm
  echo built tools/exp/lib_$v.so
done
