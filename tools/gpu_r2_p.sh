#!/bin/bash
mkdir -p gpurun_out
PLL_B200_LIB=tools/exp/lib_TIMING.so PLL_GPU_FUSED_AA=2 timeout -s KILL 90 python tools/quick_bench.py --states 20 --tips 500 --sites 200000 --iters 3 2>&1 | tail -12 > gpurun_out/p_timing.txt
cat gpurun_out/p_timing.txt
