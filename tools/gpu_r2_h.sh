#!/bin/bash
mkdir -p gpurun_out
{ ./tools/newton_c 64 1000000; PLL_GPU_L2_PERSIST=0 ./tools/newton_c 64 1000000; ./tools/newton_c 64 1000000 3; PLL_GPU_HOST_REDUCE=1 ./tools/newton_c 64 1000000 3; } 2>&1 | tee gpurun_out/h_newton_c.txt
timeout 300 python tools/newton_bench.py --tips 100 --sites 1000000 2>&1 | tail -8 | tee gpurun_out/h_newton_py.txt
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_derivatives -c 40 --csv --log-file gpurun_out/h_der_ncu.csv ./tools/newton_c 64 1000000 > /dev/null 2>&1
tail -6 gpurun_out/h_der_ncu.csv
