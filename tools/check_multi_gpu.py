#!/usr/bin/env python
"""torchrun check of the N>1 path on real GPUs (NCCL):
     python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_multi_gpu.py
Every rank holds a shard of a 2-D MFCC batch; with `process_group` set the batch-global top_db clamp
must reproduce what ONE GPU computes on the whole batch (reference functional.py:395-399)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import audio_b200.transforms as T  # noqa: E402
from audio_b200._bookkeeping import shard_bounds  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(16, 48000, generator=g)
    x[8:] *= 1e-3  # the upper shards are 60 dB quieter: they sit under the batch-global floor
    lo, hi = shard_bounds(x.shape[0], world, rank)
    mf = T.MFCC(16000, n_mfcc=40, melkwargs=dict(n_fft=1024, hop_length=256, n_mels=80)).to(dev)
    with torch.inference_mode():
        full = mf(x.to(dev))  # whole batch on this GPU, no group
        mf.process_group = dist.group.WORLD
        mine = mf(x[lo:hi].to(dev))  # this rank's shard, one all-reduce(MAX) inside
        mf.process_group = None
        local_only = mf(x[lo:hi].to(dev))
    ok = torch.allclose(mine, full[lo:hi], atol=1e-4)
    differs = not torch.allclose(local_only, full[lo:hi], atol=1e-2)
    flags = torch.tensor([int(ok), int(differs)], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN if False else dist.ReduceOp.SUM)
    if rank == 0:
        print(f"world={world} sharded==full on {flags[0].item()}/{world} ranks; "
              f"un-reduced differs on {flags[1].item()} ranks (expected >= 1)")
        assert flags[0].item() == world and flags[1].item() >= 1
        print("multi-gpu MFCC check ok")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
