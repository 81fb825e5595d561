#!/bin/bash
mkdir -p gpurun_out /tmp/ncu
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/pytest_g.txt
timeout 300 python tools/ab_time.py - > gpurun_out/ab_g.txt 2>&1
python tools/rs_bench.py > gpurun_out/rs_g.txt 2>&1
prof() {  # name, kernel regex, script
  ncu --set full --clock-control none --import-source on -k regex:$2 -s 2 -c 1 -o /tmp/ncu/$1 python $3 > gpurun_out/ncu_$1.log 2>&1
  python profiles/summarize_ncu.py /tmp/ncu/$1.ncu-rep 100 > gpurun_out/$1.txt 2>&1
  ncu -i /tmp/ncu/$1.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/$1.source.csv.gz
}
prof r2_rs_v5 resample_mma tools/rs_bench.py
prof r2_mel_v3 stft_pow2_mel tools/prof_c2.py
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err
cat gpurun_out/pytest_g.txt gpurun_out/ab_g.txt gpurun_out/rs_g.txt; head -16 gpurun_out/r2_rs_v5.txt; tail -3 gpurun_out/bench_g.err
