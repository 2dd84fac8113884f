#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sharded_nccl_gpu.py tests/test_device_slices_gpu.py -m gpu -q 2>&1 | tail -6 | tee gpurun_out/k_pytest.txt
{ ./tools/newton_c 64 2000000 2; PLL_GPU_DEVICE_REDUCE=1 ./tools/newton_c 64 2000000 2; } 2>&1 | tee gpurun_out/k_newton_c.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/k_bench_2gpu.json 2> gpurun_out/k_bench_2gpu.log
echo "bench rc=$?"; tail -3 gpurun_out/k_bench_2gpu.log; python -c "
import json; d=json.load(open('gpurun_out/k_bench_2gpu.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['config']['sharding']); print(json.dumps(d['also'])[:1500])"
