#!/bin/bash
# compute-sanitizer memcheck over the code added at the end of round 2: the pll_core_* surface (scratch
# partitions, plg_set_invariant / plg_set_sumtable) and the root log-likelihood of sliced per-rate partitions
# (plg_root_loglikelihood_counts)
mkdir -p gpurun_out
{
echo "## memcheck: pytest tests/test_core_api_gpu.py (pll_core_* through scratch partitions: DNA fused kernel, DMMA kernels, generic kernels)"
timeout -s KILL 120 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_core_api_gpu.py -x -q -m gpu 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|out of bounds" | head -12
echo "## memcheck: pytest tests/test_device_slices_gpu.py -k per_rate (plg_root_loglikelihood_counts)"
timeout -s KILL 60 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_device_slices_gpu.py -x -q -m gpu -k "per_rate_scalers" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|out of bounds" | head -12
} > gpurun_out/san2.txt 2>&1
cat gpurun_out/san2.txt
