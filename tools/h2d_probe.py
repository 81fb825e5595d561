#!/usr/bin/env python
"""Where does the 8-GPU end-to-end number go?  Under torchrun (one rank per GPU): pinned-host <-> device bandwidth per
rank, alone and with every rank copying at once, and the HostPipeline step time per rank.  B200A_NO_NUMA=1 skips the
NUMA binding for comparison.
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/h2d_probe.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_b200 import _numa  # noqa: E402
import audio_b200.transforms as T  # noqa: E402
from audio_b200.pipeline import HostPipeline  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
numa = {"node": None} if os.environ.get("B200A_NO_NUMA") else _numa.bind_to_gpu(local)
dist.init_process_group("nccl", device_id=dev)

B, L = 256, 160000
src = torch.randn(B, L).pin_memory()
dst_dev = torch.empty(B, L, device=dev)
out_dev = torch.randn(B, 626, 80, device=dev)
out_host = torch.empty(B, 626, 80).pin_memory()
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    s_in.synchronize(); s_out.synchronize()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def h2d():
    with torch.cuda.stream(s_in):
        dst_dev.copy_(src, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s_in)


def both():
    with torch.cuda.stream(s_in):
        dst_dev.copy_(src, non_blocking=True)
    with torch.cuda.stream(s_out):
        out_host.copy_(out_dev, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s_in)
    torch.cuda.current_stream().wait_stream(s_out)


mel = T.MelSpectrogram(16000, n_fft=1024, hop_length=256, n_mels=80).to(dev)
pipes = {c: HostPipeline(mel, chunk_rows=c) for c in (32, 64, 128, 256)}


def make_e2e(c):
    def e2e():
        pipes[c](src, out_host)  # result buffer re-used: no pinned allocation inside the timed region
        pipes[c].join()
    return e2e


res = {"rank": rank, "numa": numa}
res["h2d_all_ms"] = timed(h2d)
res["both_all_ms"] = timed(both)
res["e2e_all_ms"] = {c: timed(make_e2e(c)) for c in pipes}
# one rank at a time (0 and the first rank of the other socket)
def timed_solo(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    s_in.synchronize(); s_out.synchronize()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for solo in (0, world // 2):
    dist.barrier()
    if rank == solo:
        res["h2d_solo_ms"] = timed_solo(h2d)
        res["both_solo_ms"] = timed_solo(both)
        res["e2e_solo_ms"] = {c: timed_solo(make_e2e(c)) for c in pipes}
    dist.barrier()
gathered = [None] * world
dist.all_gather_object(gathered, res)
if rank == 0:
    gb = B * L * 4 / 1e6
    print(json.dumps({"no_numa": bool(os.environ.get("B200A_NO_NUMA")), "h2d_MB": gb,
                      "h2d_all_GBps": [round(gb / r["h2d_all_ms"], 1) for r in gathered],
                      "both_all_ms": [round(r["both_all_ms"], 2) for r in gathered],
                      "e2e_all_ms": {c: [round(r["e2e_all_ms"][c], 2) for r in gathered] for c in pipes},
                      "h2d_solo_GBps": {r["rank"]: round(gb / r["h2d_solo_ms"], 1) for r in gathered if "h2d_solo_ms" in r},
                      "both_solo_ms": {r["rank"]: round(r["both_solo_ms"], 2) for r in gathered if "both_solo_ms" in r},
                      "e2e_solo_ms": {r["rank"]: {c: round(v, 2) for c, v in r["e2e_solo_ms"].items()} for r in gathered if "e2e_solo_ms" in r},
                      "numa": [r["numa"].get("node") for r in gathered]}))
dist.destroy_process_group()
