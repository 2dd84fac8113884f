#!/bin/bash
# 8 GPUs, one rank each: C2 weak scaling + C5 strong scaling with the all-reduce inside the library
mkdir -p gpurun_out
nvidia-smi -L | head -8 > gpurun_out/w_gpus.txt
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus 8 --steps 8 --warmup 3 > gpurun_out/w_bench_8gpu.json 2> gpurun_out/w_bench_8gpu.log
echo "rc=$?"; tail -3 gpurun_out/w_bench_8gpu.log
python -c "
import json; d=json.load(open('gpurun_out/w_bench_8gpu.json')); print('C2', d['n_gpus'], d['ms_per_step'], d['value']); c=d['also']['c5']; print('C5', c['ms_per_step'], c['value'], c['clocks'])"
