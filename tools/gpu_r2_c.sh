#!/bin/bash
set -x
mkdir -p gpurun_out
./tools/ubench/dmma_chains > gpurun_out/c_dmma_chains.txt 2>&1; cat gpurun_out/c_dmma_chains.txt
timeout 600 python -m pytest tests/test_fused_traversal_aa_gpu.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/c_pytest.txt
cat gpurun_out/c_pytest.txt
