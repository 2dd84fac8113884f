#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 400 python bench.py --also none --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/t_bench.json 2> gpurun_out/t_bench.log
echo rc=$?; python -c "
import json; d=json.load(open('gpurun_out/t_bench.json')); r=d['roofline']; print(d['ms_per_step'], r['avg_launch_ms'], r['frac'], r['avg_launch_ms_source'])"
