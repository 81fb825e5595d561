"""world_size-2 `gloo` test of the N>1 host logic (no GPU).

The hot path shards by utterance with no collective, except MFCC's batch-global top_db cut-off
(reference functional.py:395-399) which needs ONE all-reduce(MAX) of a scalar between the
feature kernel and the clamp+DCT kernel (audio_b200.functional.mfcc, `process_group`).
Here each rank runs the ORACLE (as a stand-in for its GPU kernels) on its shard, exchanges the maximum with the
product's own `_exchange_group_max`, and rank 0 checks the re-assembled result against the oracle on
the whole batch -- including the property that skipping the all-reduce gives a different answer.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    from audio_b200._bookkeeping import shard_bounds
    from oracle import frontend_oracle as O

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(5)
        x = torch.randn(6, 4000, generator=g)
        x[3:] *= 1e-3  # the second shard is 60 dB quieter: part of it falls under the batch-global floor
        kw = dict(n_fft=256, hop_length=64, n_mels=40)
        lo, hi = shard_bounds(x.shape[0], world, rank)
        shard = x[lo:hi].numpy()

        # stage 1 on this rank: mel -> dB (unclamped) and the local maximum
        mel = O.mel_spectrogram(shard, sample_rate=16000, **kw)
        feat = O.amplitude_to_db(mel, 10.0, 1e-10, 0.0, None)
        local_max = torch.tensor([feat.max()], dtype=torch.float32)
        # the one collective of the path -- through the product's own function (audio_b200.functional.mfcc calls it
        # between the feature kernel and the clamp + DCT kernel)
        from audio_b200.functional import _exchange_group_max

        global_max = _exchange_group_max(local_max.clone(), dist.group.WORLD)
        assert _exchange_group_max(local_max, None) is local_max  # no group: identity
        # stage 2: clamp at (global max - top_db), DCT
        dct = O.create_dct(13, 40, "ortho")

        def finish(mx):
            clamped = np.maximum(feat, float(mx) - 80.0)
            return np.swapaxes(np.swapaxes(clamped, -1, -2) @ dct, -1, -2)

        mine = torch.from_numpy(finish(global_max))
        mine_local_only = torch.from_numpy(finish(local_max))
        gathered = [torch.empty(shard_bounds(6, world, r)[1] - shard_bounds(6, world, r)[0], *mine.shape[1:],
                                dtype=mine.dtype) for r in range(world)]
        dist.all_gather(gathered, mine)
        gathered_local = [torch.empty_like(t) for t in gathered]
        dist.all_gather(gathered_local, mine_local_only)
        if rank == 0:
            full = O.mfcc(x.numpy(), 16000, 13, "ortho", False, kw)
            np.savez(out_path, sharded=torch.cat(gathered).numpy(), sharded_local=torch.cat(gathered_local).numpy(),
                     full=full, bounds=np.asarray([shard_bounds(6, world, r) for r in range(world)]))
    finally:
        dist.destroy_process_group()


def test_mfcc_global_max_protocol_world2(tmp_path):
    out = str(tmp_path / "result.npz")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    z = np.load(out)
    assert z["bounds"].tolist() == [[0, 3], [3, 6]]
    # with the all-reduce the sharded run reproduces the single-process batch exactly
    # (the maximum travels as one float32, as on the GPU path: ~1e-6 dB of rounding on the floor)
    np.testing.assert_allclose(z["sharded"], z["full"], rtol=0, atol=1e-5)
    # without it (each shard clamping at its own maximum) the quiet row comes out different
    assert np.abs(z["sharded_local"] - z["full"]).max() > 1e-3


def test_module_exposes_process_group_hook():
    sys.path.insert(0, ROOT)
    import audio_b200.transforms as T

    m = T.MFCC()
    assert hasattr(m, "process_group") and m.process_group is None
