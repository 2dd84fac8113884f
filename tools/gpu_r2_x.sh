#!/bin/bash
# two B200s: the tests that need more than one device (in-process slices across devices, device-side reduce,
# two processes with the all-reduce inside the library), and the C caller's Newton timings on two devices
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/x_gpus.txt
timeout -s KILL 600 python -m pytest tests/test_device_slices_gpu.py tests/test_sharded_nccl_gpu.py tests/test_ascbias_gpu.py -q -m gpu -rs 2>&1 | tail -12 > gpurun_out/x_pytest_2gpu.txt
cat gpurun_out/x_pytest_2gpu.txt
