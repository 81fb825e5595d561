#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/pytest_i.txt
timeout 300 python tools/ab_time.py - > gpurun_out/ab_i.txt 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_i.json 2> gpurun_out/bench_i.err
cat gpurun_out/pytest_i.txt gpurun_out/ab_i.txt; tail -3 gpurun_out/bench_i.err
