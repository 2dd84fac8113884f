#!/bin/bash
# k_traverse_dna: producer back-off A/B (nanosleep between probes of the ring), sustained C2 (30 traversals: min / median) and C5-like list
mkdir -p gpurun_out; rm -f gpurun_out/sleep_ab.txt
for round in 1 2; do
  for v in sleep64 sleep1000 sleep4000; do
    echo -n "$v (round $round): C2 min / median  " >> gpurun_out/sleep_ab.txt
    PLL_B200_LIB=tools/exp/lib_$v.so timeout -s KILL 200 python tools/quick_bench.py --states 4 --tips 1000 --sites 1000000 --iters 32 --fast-tips 2>&1 | grep "^iter" | tail -30 | awk '{print $4}' | sort -n | awk '{a[NR]=$1} END {print a[1], a[int((NR+1)/2)]}' >> gpurun_out/sleep_ab.txt
  done
done
for v in sleep64 sleep1000 sleep4000; do
  echo -n "$v: C5 slice (1.25 M patterns, 5000 taxa, recycled)  " >> gpurun_out/sleep_ab.txt
  PLL_B200_LIB=tools/exp/lib_$v.so timeout -s KILL 200 python tools/c5_slice_bench.py 2>&1 | grep "N=8" >> gpurun_out/sleep_ab.txt
done
cat gpurun_out/sleep_ab.txt
