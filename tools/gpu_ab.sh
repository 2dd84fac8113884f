#!/bin/bash
# A/B of two builds of the library on the C2 traversal, alternating to average out power-cap drift
mkdir -p gpurun_out
for round in 1 2 3; do
  for v in old new; do
    echo "== $v (round $round)"
    PLL_B200_LIB=tools/exp/lib_$v.so timeout 200 python tools/quick_bench.py --states 4 --tips 1000 --sites 1000000 --iters 6 --fast-tips 2>&1 | grep "^iter" | awk '{print $4}' | sort -n | head -3 | tr '\n' ' '
    echo
  done
done | tee gpurun_out/ab.txt
