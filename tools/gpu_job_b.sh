#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "resample or config3" 2>&1 | tail -8 > gpurun_out/pytest_b.txt
python tools/rs_bench.py > gpurun_out/rs_b.txt 2>&1
B200A_RS=mma python tools/rs_bench.py >> gpurun_out/rs_b.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:stft_pow2_power -s 2 -c 1 -o gpurun_out/r2_spec_v1 python tools/prof_spec.py > gpurun_out/ncu_spec.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stft_pow2_mel -s 2 -c 1 -o gpurun_out/r2_mel_v1 python tools/prof_c2.py > gpurun_out/ncu_mel.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:resample_simt -s 2 -c 1 -o gpurun_out/r2_rs_v1 python tools/rs_bench.py > gpurun_out/ncu_rs.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err
cat gpurun_out/pytest_b.txt gpurun_out/rs_b.txt
tail -c 3000 gpurun_out/bench_b.json; tail -5 gpurun_out/bench_b.err
ls -la gpurun_out
