#!/bin/bash
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
timeout 900 ncu --metrics $M --clock-control none --csv -c 60 --log-file gpurun_out/m_launches_c2.csv \
  python bench.py --steps 2 --warmup 1 --also none --no-cpu-baseline > gpurun_out/m_ncu_c2.log 2>&1
grep -c k_traverse_dna gpurun_out/m_launches_c2.csv
timeout 900 ncu --metrics $M --clock-control none --csv -c 400 --log-file gpurun_out/m_launches_c3.csv \
  python tools/quick_bench.py --states 20 --tips 500 --sites 200000 --iters 2 > gpurun_out/m_ncu_c3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_traverse_dna -c 1 -o gpurun_out/m_traverse_dna \
  python tools/quick_bench.py --states 4 --tips 1000 --sites 200000 --iters 1 --fast-tips > gpurun_out/m_ncu_full_dna.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_partial_dmma_aa -s 12 -c 2 -o gpurun_out/m_dmma_aa \
  python tools/quick_bench.py --states 20 --tips 500 --sites 200000 --iters 1 > gpurun_out/m_ncu_full_aa.log 2>&1
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/m_bench.json 2> gpurun_out/m_bench.log
echo "bench rc=$?"; tail -2 gpurun_out/m_bench.log
ls -la gpurun_out | grep " m_"
