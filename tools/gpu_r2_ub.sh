#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 200 tools/exp/tma_store_peak > gpurun_out/ub_tma_store_peak.txt 2>&1; cat gpurun_out/ub_tma_store_peak.txt
