#!/bin/bash
# two B200s: the tests that need more than one device (in-process slices across devices, device-side reduce,
# two processes with the all-reduce inside the library), and the C caller's Newton timings on two devices
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/x_gpus.txt
timeout -s KILL 600 python -m pytest tests/test_device_slices_gpu.py tests/test_sharded_nccl_gpu.py tests/test_ascbias_gpu.py tests/test_branch_optimisation_gpu.py tests/test_partial_traversal_gpu.py -q -m gpu -rs 2>&1 | tail -6 > gpurun_out/x_pytest_2gpu.txt
PLL_GPU_DEVICES=2 timeout -s KILL 600 python -m pytest tests/test_branch_optimisation_gpu.py tests/test_partial_traversal_gpu.py tests/test_golden_gpu.py -q -m gpu 2>&1 | tail -3 >> gpurun_out/x_pytest_2gpu.txt
cat gpurun_out/x_pytest_2gpu.txt
{ ./tools/newton_c 64 2000000 2; PLL_GPU_HOST_THREADS=0 ./tools/newton_c 64 2000000 2; } 2>&1 | grep -v "^lnL" | tee gpurun_out/x_newton_c_2gpu.txt
