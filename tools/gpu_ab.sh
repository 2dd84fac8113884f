#!/bin/bash
# A/B of two builds of the library on the C2 traversal, alternating to average out power-cap drift:
# per run the minimum and the median of 14 traversals
mkdir -p gpurun_out; rm -f gpurun_out/ab.txt
for round in 1 2 3 4; do
  for v in old new; do
    echo -n "$v (round $round): min / median  " >> gpurun_out/ab.txt
    PLL_B200_LIB=tools/exp/lib_$v.so timeout -s KILL 200 python tools/quick_bench.py --states 4 --tips 1000 --sites 1000000 --iters 16 --fast-tips 2>&1 | grep "^iter" | tail -14 | awk '{print $4}' | sort -n | awk '{a[NR]=$1} END {print a[1], a[int((NR+1)/2)]}' >> gpurun_out/ab.txt
  done
done
cat gpurun_out/ab.txt
