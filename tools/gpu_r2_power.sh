#!/bin/bash
# SM clock and board power while the C2 traversal runs back to back (40 traversals): is the sustained rate power-capped?
mkdir -p gpurun_out
nvidia-smi --query-gpu=timestamp,clocks.sm,clocks.mem,power.draw,power.limit,temperature.gpu,clocks_throttle_reasons.sw_power_cap,clocks_throttle_reasons.hw_slowdown --format=csv -lms 50 > gpurun_out/power_trace.csv &
SMI=$!
timeout -s KILL 200 python tools/quick_bench.py --states 4 --tips 1000 --sites 1000000 --iters 40 --fast-tips 2>&1 | grep "^iter" | awk '{print $2, $4}' > gpurun_out/power_iters.txt
kill $SMI
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/power_trace.csv')))[1:]
busy=[r for r in rows if float(r[3].split()[0])>400]
import statistics as st
clk=[int(r[1].split()[0]) for r in busy]; pw=[float(r[3].split()[0]) for r in busy]
print(f"samples under load {len(busy)}: SM clock min {min(clk)} median {st.median(clk)} max {max(clk)} MHz; power median {st.median(pw):.0f} W max {max(pw):.0f} W of limit {rows[0][4].strip()}; sw_power_cap active in {sum('Active' in r[6] and 'Not' not in r[6] for r in busy)} samples")
it=[float(l.split()[1]) for l in open('gpurun_out/power_iters.txt')]
print('traversal ms: first 5', it[:5], 'last 5', it[-5:])
PY
