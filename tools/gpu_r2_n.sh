#!/bin/bash
# second 20-state whole-list kernel (plg_walk_aa.cu): parity, then C3 timing against the level-by-level path
mkdir -p gpurun_out
export PLL_TEST_FUSED_AA=2
timeout -s KILL 40 python tools/quick_bench.py --states 20 --tips 40 --sites 3000 --iters 2 2>&1 | tail -2
PLL_GPU_FUSED_AA=2 timeout -s KILL 40 python tools/quick_bench.py --states 20 --tips 40 --sites 3000 --iters 2 2>&1 | tail -2 || { echo "HANG or failure in the small run"; exit 1; }
timeout -s KILL 200 python -m pytest tests/test_fused_traversal_aa_gpu.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/n_pytest.txt
cat gpurun_out/n_pytest.txt
rm -f gpurun_out/n_c3.txt
for fa in 0 2; do
  echo "== PLL_GPU_FUSED_AA=$fa" >> gpurun_out/n_c3.txt
  PLL_GPU_FUSED_AA=$fa timeout -s KILL 90 python tools/quick_bench.py --states 20 --tips 500 --sites 200000 --iters 4 2>&1 | tail -5 >> gpurun_out/n_c3.txt
done
cat gpurun_out/n_c3.txt
