#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_ascbias_gpu.py tests/test_device_slices_gpu.py -x -q -m gpu 2>&1 | tail -4 > gpurun_out/t_pytest.txt
cat gpurun_out/t_pytest.txt
{ ./tools/newton_c 64 1000000 1; ./tools/newton_c 64 1000000 3; PLL_GPU_HOST_THREADS=0 ./tools/newton_c 64 1000000 3; } 2>&1 | grep -v "^lnL" 
