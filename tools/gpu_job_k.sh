#!/bin/bash
mkdir -p gpurun_out
echo "alt (8 epilogue warps):" > gpurun_out/rs_k.txt
B200A_LIB=audio_b200/build/libb200audio_e4.so B200A_RS=tc timeout -k 10 120 python tools/rs_bench.py >> gpurun_out/rs_k.txt 2>&1
B200A_LIB=audio_b200/build/libb200audio_e4.so B200A_RS=tc timeout -k 10 120 python tools/rs_check.py cmp >> gpurun_out/rs_k.txt 2>&1
cat gpurun_out/rs_k.txt
