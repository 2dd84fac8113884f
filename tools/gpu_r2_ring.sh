#!/bin/bash
# k_traverse_dna: ring depth A/B at C2 (alternating builds to average out power-cap drift)
mkdir -p gpurun_out; rm -f gpurun_out/ring_ab.txt
for round in 1 2 3; do
  for v in S8 S12 S15; do
    echo -n "$v (round $round): " >> gpurun_out/ring_ab.txt
    PLL_B200_LIB=tools/exp/lib_$v.so timeout -s KILL 120 python tools/quick_bench.py --states 4 --tips 1000 --sites 1000000 --iters 6 --fast-tips 2>&1 | grep "^iter" | awk '{print $4}' | sort -n | head -3 | tr '\n' ' ' >> gpurun_out/ring_ab.txt
    echo >> gpurun_out/ring_ab.txt
  done
done
cat gpurun_out/ring_ab.txt
