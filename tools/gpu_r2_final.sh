#!/bin/bash
# round-2 final state: smoke, full GPU suite, default bench line, reference arm
mkdir -p gpurun_out
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.txt 2>&1; tail -1 gpurun_out/final_smoke.txt
timeout -s KILL 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4 > gpurun_out/final_pytest.txt; cat gpurun_out/final_pytest.txt
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.log
echo "bench rc=$?"; grep -v "^$" gpurun_out/final_bench.log | tail -14
python -c "
import json; d=json.load(open('gpurun_out/final_bench.json')); print('C2', d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'], 'lnl ok', d['lnl_check']['ok']); a=d['also']; print('C3', a['c3']['ms_per_step'], 'C4', a['c4']['value'], a['c4']['derivative_call_us'], a['c4']['c_caller'], 'C5', a['c5']['ms_per_step'], 'lists', a['lists']['host_overhead_frac_of_device'])"
timeout -s KILL 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.log; tail -c 600 gpurun_out/final_bench_reference.json
