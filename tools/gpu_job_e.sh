#!/bin/bash
mkdir -p gpurun_out
python tools/ab_time.py - B200A_LIB=audio_b200/build/libb200audio_w0.so - B200A_LIB=audio_b200/build/libb200audio_w0.so > gpurun_out/ab_e.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/pytest_e.txt
cat gpurun_out/ab_e.txt gpurun_out/pytest_e.txt
