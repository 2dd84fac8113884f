#!/bin/bash
# determinism of the walk kernel: the same traversal many times, lnL must not move
mkdir -p gpurun_out; rm -f gpurun_out/v_determinism.txt
for v in head early; do
  for sites in 20000 200000; do
    echo "== $v sites=$sites" >> gpurun_out/v_determinism.txt
    it=300; [ $sites = 200000 ] && it=60
    PLL_B200_LIB=tools/exp/lib_$v.so PLL_GPU_FUSED_AA=1 timeout -s KILL 200 python tools/quick_bench.py --states 20 --tips 500 --sites $sites --iters $it 2>&1 | grep -o "lnL=.*" | sort | uniq -c >> gpurun_out/v_determinism.txt
  done
done
cat gpurun_out/v_determinism.txt
