#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/ab_time.py - B200A_LIB=audio_b200/build/libb200audio_n12.so B200A_LIB=audio_b200/build/libb200audio_tw8.so - B200A_LIB=audio_b200/build/libb200audio_n12.so B200A_TC=0 > gpurun_out/ab_f.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/pytest_f.txt
cut -c1-300 gpurun_out/ab_f.txt; cat gpurun_out/pytest_f.txt
