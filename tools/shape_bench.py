#!/usr/bin/env python
"""Per-frame cost of the fused kernels as the clip length changes (same total samples): short clips have a larger
share of edge units (reflect padding at both ends), which take the gather path instead of the bulk-staged one.
    python tools/shape_bench.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import audio_b200.transforms as T  # noqa: E402
from tools.bench_configs import time_gpu  # noqa: E402


def main():
    mods = (("MelSpectrogram 1024/256/80", lambda: T.MelSpectrogram(16000, n_fft=1024, hop_length=256, n_mels=80)),
            ("Spectrogram 1024/256", lambda: T.Spectrogram(n_fft=1024, hop_length=256)),
            ("MFCC 40 over mel 1024/256/80", lambda: T.MFCC(16000, n_mfcc=40, melkwargs=dict(n_fft=1024, hop_length=256, n_mels=80))))
    for rows, length in ((4, 10240000), (256, 160000), (2048, 20000), (8192, 5000)):
        x = torch.randn(rows, length, device="cuda")
        for name, make in mods:
            m = make().cuda()
            y = m(x)
            torch.cuda.synchronize()
            frames = y.shape[0] * y.shape[-1]
            t = time_gpu(lambda: m(x))
            print(f"{name:32s} {rows:5d} x {length:8d}: {t:.4f} ms  {t * 1e6 / frames:.3f} ns/frame", flush=True)
            del y


if __name__ == "__main__":
    main()
