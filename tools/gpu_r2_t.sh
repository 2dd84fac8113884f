#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/t_pytest.txt
cat gpurun_out/t_pytest.txt
