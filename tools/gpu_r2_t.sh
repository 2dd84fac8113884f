#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.log
echo rc=$?; python -c "
import json; d=json.load(open('gpurun_out/final_bench.json')); r=d['roofline']; print('C2', d['ms_per_step'], r['avg_launch_ms'], r['frac'], d['clocks']); a=d['also']; print('C3', a['c3']['ms_per_step'], a['c3']['recycled_slots']['default']['ms_per_step'], 'C4', a['c4']['value'], 'C5', a['c5']['ms_per_step'])"
