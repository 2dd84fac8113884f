#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/g_pytest.txt
cat gpurun_out/g_pytest.txt
timeout 300 python tools/newton_bench.py --tips 100 --sites 1000000 2>&1 | tail -8 | tee gpurun_out/g_newton.txt
timeout 600 python tools/inprocess_devices_bench.py --devices 3 --sites-per-device 200000 --tips 200 --steps 5 2>&1 | tail -4 | tee gpurun_out/g_inproc_1gpu_3slices.txt
