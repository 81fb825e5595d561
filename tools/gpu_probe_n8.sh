#!/bin/bash
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tools/h2d_probe.py > gpurun_out/h2d_n8.txt 2> gpurun_out/h2d_n8.err
grep "^{" gpurun_out/h2d_n8.txt; tail -3 gpurun_out/h2d_n8.err; true
