#!/usr/bin/env python
"""Kaldi-compatible fbank / mfcc throughput on one GPU (the ASR front-end configuration: 16 kHz, 25 ms / 10 ms
frames = 400-sample povey window in a 512-point FFT, 80 mel bins), batch 256 x 10 s, vs torchaudio CPU.
    python tools/kaldi_bench.py
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import audio_b200.compliance.kaldi as K  # noqa: E402
from tools.bench_configs import time_gpu  # noqa: E402


def main():
    x = (torch.randn(256, 160000, device="cuda") * 3000.0).round()
    for name, fn in (("fbank 80 bins", lambda: K.fbank_batch(x, num_mel_bins=80)),
                     ("fbank 80 bins + energy, snip_edges=False", lambda: K.fbank_batch(x, num_mel_bins=80, use_energy=True, snip_edges=False)),
                     ("mfcc 13 of 23", lambda: K.mfcc_batch(x)),
                     ("spectrogram (257 log-power bins)", lambda: K.spectrogram_batch(x))):
        y = fn()
        torch.cuda.synchronize()
        frames = y.shape[0] * y.shape[1]
        t = time_gpu(fn, iters=10, blocks=3)
        print(f"{name:45s} {t:8.4f} ms  {frames / (t * 1e-3):14.4g} frames/s", flush=True)
    try:
        import torchaudio

        xc = x[:16, :].cpu()
        t0 = time.perf_counter()
        for r in range(16):
            torchaudio.compliance.kaldi.fbank(xc[r:r + 1], num_mel_bins=80)
        dt = time.perf_counter() - t0
        print(f"torchaudio CPU fbank 80 bins: {16 * 998 / dt:12.0f} frames/s ({torch.get_num_threads()} threads)")
    except Exception as e:  # noqa: BLE001
        print("torchaudio CPU reference unavailable:", e)


if __name__ == "__main__":
    main()
