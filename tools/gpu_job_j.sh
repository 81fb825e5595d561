#!/bin/bash
mkdir -p gpurun_out /tmp/ncu
B200A_RS=mma timeout -k 10 120 python tools/rs_check.py save > gpurun_out/rs_j.txt 2>&1; bash tools/gpu_job_k.sh
B200A_RS=tc timeout -k 10 120 python tools/rs_check.py cmp >> gpurun_out/rs_j.txt 2>&1
echo "rc=$?" >> gpurun_out/rs_j.txt
B200A_RS=tc timeout -k 10 120 python tools/rs_bench.py >> gpurun_out/rs_j.txt 2>&1
B200A_RS=mma timeout -k 10 120 python tools/rs_bench.py >> gpurun_out/rs_j.txt 2>&1
B200A_RS=tc timeout 300 ncu --set full --clock-control none --import-source on -k regex:resample_tc_kernel -s 2 -c 1 -o /tmp/ncu/r2_rs_tc python tools/rs_bench.py > gpurun_out/ncu_rs_tc.log 2>&1
python profiles/summarize_ncu.py /tmp/ncu/r2_rs_tc.ncu-rep 60 > gpurun_out/r2_rs_tc.txt 2>&1
cat gpurun_out/rs_j.txt; head -60 gpurun_out/r2_rs_tc.txt
