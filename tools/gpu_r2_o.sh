#!/bin/bash
mkdir -p gpurun_out
export PLL_GPU_FUSED_AA=2
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_walk_aa -c 1 -o gpurun_out/o_walk_aa \
  python tools/quick_bench.py --states 20 --tips 500 --sites 100000 --iters 1 > gpurun_out/o_ncu.log 2>&1
tail -3 gpurun_out/o_ncu.log
