#!/bin/bash
# round-2 GPU session A: full GPU suite, C3 timing fused vs level-by-level, ncu of k_traverse_aa
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/a_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/a_pytest.txt
cat gpurun_out/a_pytest.txt
for f in 1 0; do
  PLL_GPU_FUSED=$f timeout 300 python tools/quick_bench.py --states 20 --tips 500 --sites 200000 --iters 6 > gpurun_out/a_c3_fused$f.txt 2>&1
  tail -7 gpurun_out/a_c3_fused$f.txt
done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  --log-file gpurun_out/a_launches_c3_fused.csv python tools/quick_bench.py --states 20 --tips 500 --sites 200000 --iters 2 > gpurun_out/a_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_traverse_aa -c 1 -o gpurun_out/a_traverse_aa \
  python tools/quick_bench.py --states 20 --tips 500 --sites 200000 --iters 1 > gpurun_out/a_ncu2.log 2>&1
ls -la gpurun_out
