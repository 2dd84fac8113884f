#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_fused_traversal_aa_gpu.py -x -q -m gpu -k "benchmark_shape or deterministic" 2>&1 | tail -6 > gpurun_out/t_pytest.txt
cat gpurun_out/t_pytest.txt
