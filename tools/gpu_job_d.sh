#!/bin/bash
mkdir -p gpurun_out
python tools/skew_sweep.py 0 800 1600 2400 3200 4800 > gpurun_out/skew_d.txt 2>&1
python tools/rs_bench.py > gpurun_out/rs_d.txt 2>&1
B200A_RS=mma python tools/rs_bench.py >> gpurun_out/rs_d.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q -k "resample or config3 or pipeline or plan or exchange" 2>&1 | tail -6 > gpurun_out/pytest_d.txt
cat gpurun_out/skew_d.txt gpurun_out/rs_d.txt gpurun_out/pytest_d.txt
