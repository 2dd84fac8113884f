#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 400 python bench.py --also lists,c4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/y_bench_c4.json 2> gpurun_out/y_bench_c4.log
grep "c4 branch" gpurun_out/y_bench_c4.log
python -c "
import json; d=json.load(open('gpurun_out/y_bench_c4.json')); c=d['also']['c4']; print({k:c[k] for k in ('sumtable_call_us','derivative_call_us','reroot_call_us','value')})"
