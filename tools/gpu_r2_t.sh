#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 400 python bench.py --also c3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/t_bench_c3.json 2> gpurun_out/t_bench_c3.log
echo rc=$?; tail -2 gpurun_out/t_bench_c3.log; python -c "
import json; d=json.load(open('gpurun_out/t_bench_c3.json')); c=d['also']['c3']; print(c['ms_per_step'], json.dumps(c['recycled_slots'], indent=1))"
