#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fused_traversal_gpu.py tests/test_parity_gpu.py tests/test_partial_traversal_gpu.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/l_pytest.txt
timeout 300 python tools/quick_bench.py --states 4 --tips 1000 --sites 1000000 --iters 5 --fast-tips 2>&1 | tail -3 | tee gpurun_out/l_c2.txt
