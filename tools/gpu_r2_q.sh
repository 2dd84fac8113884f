#!/bin/bash
# level-by-level 20-state kernel with two bulk copies per child and unit: parity subset + C3 timing per kind
mkdir -p gpurun_out
timeout -s KILL 60 python tools/quick_bench.py --states 20 --tips 40 --sites 3000 --iters 2 2>&1 | tail -1 || exit 1
timeout -s KILL 500 python -m pytest tests/test_parity_gpu.py tests/test_fused_traversal_aa_gpu.py tests/test_lg4_example_gpu.py tests/test_golden_gpu.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/q_pytest.txt
cat gpurun_out/q_pytest.txt
PLL_GPU_FUSED_AA=0 timeout -s KILL 90 python tools/quick_bench.py --states 20 --tips 500 --sites 200000 --iters 5 2>&1 | tail -4 > gpurun_out/q_c3.txt
timeout -s KILL 120 python tools/kind_bench.py --states 20 --tips 500 --sites 200000 2>&1 | tail -8 >> gpurun_out/q_c3.txt
cat gpurun_out/q_c3.txt
