#!/bin/bash
# full GPU suite, default bench line, C5 slice scaling on one GPU
mkdir -p gpurun_out
timeout -s KILL 60 python tools/quick_bench.py --states 20 --tips 40 --sites 3000 --iters 2 2>&1 | tail -1 || exit 1
timeout -s KILL 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/r_pytest.txt
cat gpurun_out/r_pytest.txt
timeout -s KILL 300 python tools/c5_slice_bench.py > gpurun_out/r_c5_slices.txt 2>&1; cat gpurun_out/r_c5_slices.txt
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r_bench.json 2> gpurun_out/r_bench.log
echo "bench rc=$?"; tail -2 gpurun_out/r_bench.log; python -c "
import json; d=json.load(open('gpurun_out/r_bench.json')); print(d['ms_per_step'], d['roofline']['frac'], d['also']['c3']['ms_per_step'], d['also']['c5']['ms_per_step'], d['also']['c4']['derivative_call_us'])"
