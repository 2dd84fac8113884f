#!/bin/bash
mkdir -p gpurun_out
{ ./tools/newton_c 64 1000000; PLL_GPU_L2_PERSIST=0 ./tools/newton_c 64 1000000; } 2>&1 | tee gpurun_out/i_newton_c.txt
timeout 300 python tools/quick_bench.py --states 4 --tips 1000 --sites 1000000 --iters 4 --fast-tips 2>&1 | tail -2 | tee gpurun_out/i_c2.txt
