#!/bin/bash
# tests + bench + ncu summaries (text only: the .ncu-rep files are ~37 MB each and gpurun_out is capped at 64 MiB)
mkdir -p gpurun_out /tmp/ncu
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_c.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err
prof() {  # name, kernel regex, script
  ncu --set full --clock-control none --import-source on -k regex:$2 -s 2 -c 1 -o /tmp/ncu/$1 python $3 > gpurun_out/ncu_$1.log 2>&1
  python profiles/summarize_ncu.py /tmp/ncu/$1.ncu-rep 100 > gpurun_out/$1.txt 2>&1
  ncu -i /tmp/ncu/$1.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/$1.source.csv.gz
}
prof r2_mel_v2 stft_pow2_mel tools/prof_c2.py
prof r2_spec_v1 stft_pow2_power tools/prof_spec.py
cat gpurun_out/pytest_c.txt; tail -c 2500 gpurun_out/bench_c.json; tail -5 gpurun_out/bench_c.err; head -20 gpurun_out/r2_mel_v2.txt
ls -la gpurun_out
