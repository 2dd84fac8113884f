#!/bin/bash
# ONE process, ONE pll_partition_t over 8 B200s (pll_gpu_set_devices): C caller's call latencies, 1 M patterns per device;
# then the in-process multi-device tests on 8 devices
mkdir -p gpurun_out
{ echo "## 1 device, 1 M patterns"; ./tools/newton_c 64 1000000 1
  echo "## 8 devices in one process, 8 M patterns: helper threads (default), host sum of the per-device results"; ./tools/newton_c 64 8000000 8
  echo "## the same, launches from the calling thread only (PLL_GPU_HOST_THREADS=0)"; PLL_GPU_HOST_THREADS=0 ./tools/newton_c 64 8000000 8
  echo "## the same, PLL_GPU_DEVICE_REDUCE=1 (devices combine through peer-mapped slots; calls stay on one thread)"; PLL_GPU_DEVICE_REDUCE=1 ./tools/newton_c 64 8000000 8
  echo "## 8 devices, 1 M patterns in total (strong scaling of a small problem)"; ./tools/newton_c 64 1000000 8
} > gpurun_out/dev8_newton_c.txt 2>&1
cat gpurun_out/dev8_newton_c.txt
PLL_GPU_DEVICES=8 timeout -s KILL 300 python -m pytest tests/test_device_slices_gpu.py tests/test_ascbias_gpu.py -q -m gpu 2>&1 | tail -3 | tee gpurun_out/dev8_pytest.txt
