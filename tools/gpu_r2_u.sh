#!/bin/bash
# compute-sanitizer memcheck over the round-2 kernels: the 20-state walk (k_walk_aa, k_walk_pack_aa), the
# level-by-level DMMA kernels with the two-half ring, sliced ascertainment-bias partitions
mkdir -p gpurun_out
{
echo "## memcheck: pytest tests/test_fused_traversal_aa_gpu.py -k 'equals_level_by_level and 40-1000' (k_walk_aa, k_walk_pack_aa, k_partial_dmma_aa, k_partial_tt_aa)"
PLL_GPU_GRAPHS=0 timeout -s KILL 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_fused_traversal_aa_gpu.py -x -q -m gpu -k "equals_level_by_level and 40-1000" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|out of bounds" | head -12
echo "## memcheck: pytest tests/test_fused_traversal_aa_gpu.py -k 'recycled or lg4m' (tile-cache misses, dead stores, rescaling fix-up)"
PLL_GPU_GRAPHS=0 timeout -s KILL 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_fused_traversal_aa_gpu.py -x -q -m gpu -k "recycled or lg4m" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|out of bounds" | head -12
echo "## memcheck: pytest tests/test_ascbias_gpu.py (sliced partitions with the correction, zero-pattern slices)"
timeout -s KILL 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_ascbias_gpu.py -x -q -m gpu 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|out of bounds" | head -12
} > gpurun_out/u_sanitizer.txt 2>&1
cat gpurun_out/u_sanitizer.txt
