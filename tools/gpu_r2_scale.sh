#!/bin/bash
# torchrun on N GPUs of one box: C2 weak scaling + C5 strong scaling (N = number of visible GPUs)
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 \
  bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/scale_bench_${N}gpu.json 2> gpurun_out/scale_bench_${N}gpu.log
echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/scale_bench_${N}gpu.json')); print('C2', d['n_gpus'], d['ms_per_step'], d['value'], d['clocks']); c=d['also']['c5']; print('C5', c['ms_per_step'], c['value'], c['lnl'])"
