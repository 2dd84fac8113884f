#!/bin/bash
# profiles of the shipping 20-state kernels after the ring change, and of the opt-in walk kernel
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
timeout -s KILL 400 ncu --metrics $M --clock-control none --csv -c 400 --log-file gpurun_out/s_launches_c3.csv \
  python tools/quick_bench.py --states 20 --tips 500 --sites 200000 --iters 2 > gpurun_out/s_ncu_c3.log 2>&1
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:k_partial_dmma_aa -s 12 -c 2 -o gpurun_out/s_dmma_aa \
  python tools/quick_bench.py --states 20 --tips 500 --sites 200000 --iters 1 > gpurun_out/s_ncu_full_aa.log 2>&1
PLL_GPU_FUSED_AA=1 timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:k_walk_aa -c 1 -o gpurun_out/s_walk_aa \
  python tools/quick_bench.py --states 20 --tips 500 --sites 200000 --iters 1 > gpurun_out/s_ncu_full_walk.log 2>&1
PLL_GPU_FUSED_AA=1 timeout -s KILL 400 ncu --metrics $M --clock-control none --csv -c 20 --log-file gpurun_out/s_launches_c3_walk.csv \
  python tools/quick_bench.py --states 20 --tips 500 --sites 200000 --iters 2 > gpurun_out/s_ncu_c3_walk.log 2>&1
ls -la gpurun_out | grep " s_"
