#!/bin/bash
# DNA traversal A/B (old = committed build, new = working tree): parity of the new build first
mkdir -p gpurun_out
timeout -s KILL 60 python tools/quick_bench.py --states 4 --tips 40 --sites 3000 --iters 2 2>&1 | tail -1 || exit 1
timeout -s KILL 900 python -m pytest tests/test_fused_traversal_gpu.py tests/test_parity_gpu.py tests/test_golden_gpu.py tests/test_partial_traversal_gpu.py tests/test_synthetic_tips_gpu.py -x -q -m gpu 2>&1 | tail -4 > gpurun_out/z_pytest.txt
cat gpurun_out/z_pytest.txt
rm -f gpurun_out/z_c2.txt
for round in 1 2 3; do
  for v in old new; do
    echo -n "$v (round $round): " >> gpurun_out/z_c2.txt
    PLL_B200_LIB=tools/exp/lib_$v.so timeout -s KILL 120 python tools/quick_bench.py --states 4 --tips 1000 --sites 1000000 --iters 6 --fast-tips 2>&1 | grep "^iter" | awk '{print $4}' | sort -n | head -3 | tr '\n' ' ' >> gpurun_out/z_c2.txt
    echo >> gpurun_out/z_c2.txt
  done
done
cat gpurun_out/z_c2.txt
