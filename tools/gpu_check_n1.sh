#!/bin/bash
# final check of the round: full GPU test suite, the driver-style bench line, one ncu capture of the tcgen05 resampler
mkdir -p gpurun_out /tmp/ncu
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/pytest_m.txt
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_m.json 2> gpurun_out/bench_m.err
timeout 200 ncu --set full --clock-control none --import-source on -k regex:resample_tc_kernel -s 2 -c 1 -o /tmp/ncu/r2_rs_tc python tools/rs_bench.py > gpurun_out/ncu_rs_tc.log 2>&1
python profiles/summarize_ncu.py /tmp/ncu/r2_rs_tc.ncu-rep 60 > gpurun_out/r2_resample_tc.txt 2>&1
cat gpurun_out/pytest_m.txt; cut -c1-300 gpurun_out/bench_m.json; tail -3 gpurun_out/bench_m.err; head -20 gpurun_out/r2_resample_tc.txt
