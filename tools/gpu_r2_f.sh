#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_traverse_dna -c 1 -o gpurun_out/f_traverse_dna \
  python tools/quick_bench.py --states 4 --tips 1000 --sites 200000 --iters 1 --fast-tips > gpurun_out/f_ncu.log 2>&1
tail -3 gpurun_out/f_ncu.log
timeout 300 python tools/quick_bench.py --states 4 --tips 1000 --sites 1000000 --iters 4 --fast-tips 2>&1 | tail -3 | tee gpurun_out/f_c2_quick.txt
