#!/bin/bash
# round-2 GPU session B: AA fused tests, C3 timing, ncu of the reworked k_traverse_aa
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fused_traversal_aa_gpu.py tests/test_synthetic_tips_gpu.py tests/test_parity_gpu.py tests/test_lg4_example_gpu.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/b_pytest.txt
cat gpurun_out/b_pytest.txt
PLL_GPU_FUSED=1 timeout 300 python tools/quick_bench.py --states 20 --tips 500 --sites 200000 --iters 6 > gpurun_out/b_c3_fused1.txt 2>&1
tail -7 gpurun_out/b_c3_fused1.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_traverse_aa -c 1 -o gpurun_out/b_traverse_aa \
  python tools/quick_bench.py --states 20 --tips 500 --sites 200000 --iters 1 > gpurun_out/b_ncu2.log 2>&1
ls -la gpurun_out | tail -5
