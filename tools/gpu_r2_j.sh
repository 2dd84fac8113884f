#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/j_gpus.txt
timeout 600 python -m pytest tests/test_device_slices_gpu.py -m gpu -q 2>&1 | tail -4 | tee gpurun_out/j_pytest.txt
{ ./tools/newton_c 64 1000000 1; ./tools/newton_c 64 2000000 2; PLL_GPU_HOST_REDUCE=1 ./tools/newton_c 64 2000000 2; } 2>&1 | tee gpurun_out/j_newton_c.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/j_bench_2gpu.json 2> gpurun_out/j_bench_2gpu.log
echo "bench rc=$?"; tail -3 gpurun_out/j_bench_2gpu.log; head -c 600 gpurun_out/j_bench_2gpu.json
