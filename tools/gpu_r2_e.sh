#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fused_traversal_aa_gpu.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/e_pytest.txt
cat gpurun_out/e_pytest.txt
timeout 200 python tools/quick_bench.py --states 20 --tips 500 --sites 200000 --iters 5 --fast-tips 2>&1 | tail -3 | tee gpurun_out/e_c3.txt
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/e_bench.json 2> gpurun_out/e_bench.log
echo "bench rc=$?"; tail -5 gpurun_out/e_bench.log; head -c 3000 gpurun_out/e_bench.json
