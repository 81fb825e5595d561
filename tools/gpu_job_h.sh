#!/bin/bash
mkdir -p gpurun_out /tmp/ncu
timeout 120 python tools/rs_bench.py > gpurun_out/rs_h.txt 2>&1
B200A_RS=mma timeout 120 python tools/rs_bench.py >> gpurun_out/rs_h.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q -k "resample or config3" 2>&1 | tail -8 > gpurun_out/pytest_h.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:resample_r3 -s 2 -c 1 -o /tmp/ncu/r2_rs_r3 python tools/rs_bench.py > gpurun_out/ncu_rs_r3.log 2>&1
python profiles/summarize_ncu.py /tmp/ncu/r2_rs_r3.ncu-rep 60 > gpurun_out/r2_rs_r3.txt 2>&1
cat gpurun_out/rs_h.txt gpurun_out/pytest_h.txt; head -40 gpurun_out/r2_rs_r3.txt
