#!/bin/bash
# 8-GPU check: NUMA-bound host pipeline (e2e scaling), C4 with the NCCL all-reduce live, sharded-MFCC parity test
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n8.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
timeout 300 python -m pytest tests -m gpu -q -k "sharded_mfcc" 2>&1 | tail -4 > gpurun_out/pytest_n8.txt
python - <<'PY'
import json
for f in ("gpurun_out/bench_n8.json",):
    try:
        b = json.load(open(f))
        print(f, "value", b["value"], "ms", b["ms_per_step"], "e2e", b["e2e"]["value"], b["e2e"]["ms_per_step"], b["e2e"].get("numa"))
        for c in b["configs"]:
            print("  ", c["key"], round(c["ms_per_step"], 4), c.get("collective_cost_ms"))
    except Exception as e:
        print(f, "failed", e)
PY
cat gpurun_out/pytest_n8.txt; tail -5 gpurun_out/bench_n8.err
