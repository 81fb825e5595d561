#!/bin/bash
mkdir -p gpurun_out /tmp/ncu
B200A_DEBUG=1 timeout 120 python tools/rs_bench.py > gpurun_out/rs_h.txt 2>&1
B200A_RS=mma timeout 120 python tools/rs_bench.py >> gpurun_out/rs_h.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/pytest_h.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:resample_r3_kernel -s 2 -c 1 -o /tmp/ncu/r2_rs_r3 python tools/rs_bench.py > gpurun_out/ncu_rs_r3.log 2>&1
python profiles/summarize_ncu.py /tmp/ncu/r2_rs_r3.ncu-rep 60 > gpurun_out/r2_rs_r3.txt 2>&1
timeout 200 python tools/ab_time.py - > gpurun_out/ab_h.txt 2>&1
sort gpurun_out/rs_h.txt | uniq -c | sort -rn | head -12; cat gpurun_out/pytest_h.txt gpurun_out/ab_h.txt; head -40 gpurun_out/r2_rs_r3.txt
