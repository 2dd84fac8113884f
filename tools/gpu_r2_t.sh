#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_ascbias_gpu.py tests/test_device_slices_gpu.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/t_pytest.txt
cat gpurun_out/t_pytest.txt
