#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_l.txt
timeout 120 python tools/rs_bench.py > gpurun_out/rs_l.txt 2>&1
timeout 300 python tools/ab_time.py - >> gpurun_out/rs_l.txt 2>&1
cat gpurun_out/pytest_l.txt gpurun_out/rs_l.txt
