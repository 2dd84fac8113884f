#!/bin/bash
# DNA traversal with the pair table of tip-tip operations: parity, then C2 timing (A/B against the previous build)
mkdir -p gpurun_out
timeout -s KILL 60 python tools/quick_bench.py --states 4 --tips 40 --sites 3000 --iters 2 2>&1 | tail -1 || exit 1
timeout -s KILL 600 python -m pytest tests/test_fused_traversal_gpu.py tests/test_parity_gpu.py tests/test_golden_gpu.py tests/test_partial_traversal_gpu.py -x -q -m gpu 2>&1 | tail -4 > gpurun_out/z_pytest.txt
cat gpurun_out/z_pytest.txt
rm -f gpurun_out/z_c2.txt
for round in 1 2 3; do
  for v in old new; do
    lib=""; [ $v = old ] && lib="tools/exp/lib_old.so"
    echo "== $v (round $round)" >> gpurun_out/z_c2.txt
    PLL_B200_LIB=$lib timeout -s KILL 120 python tools/quick_bench.py --states 4 --tips 1000 --sites 1000000 --iters 5 --fast-tips 2>&1 | grep -o "traversal [0-9.]* ms" | tail -3 | tr '\n' ' ' >> gpurun_out/z_c2.txt
    echo >> gpurun_out/z_c2.txt
  done
done
cat gpurun_out/z_c2.txt
