#!/bin/bash
mkdir -p gpurun_out
for v in "" NOSTG NOVOTE NOCACHE NOTABLE NOMMA NOSTG+NOVOTE+NOCACHE+NOTABLE; do
  if [ -z "$v" ]; then unset PLL_B200_LIB; else export PLL_B200_LIB=tools/exp/lib_$v.so; fi
  echo "== variant [$v]"
  timeout 200 python tools/quick_bench.py --states 20 --tips 500 --sites 200000 --iters 4 --fast-tips 2>&1 | tail -2
done > gpurun_out/d_variants.txt 2>&1
cat gpurun_out/d_variants.txt
