#!/usr/bin/env python
"""Timing of the paths that run on the generic (any-size, shared-memory Stockham) kernels: the reference's DEFAULT
n_fft = 400 transforms, complex STFT, inverse STFT, Griffin-Lim, pitch shift.
    python tools/generic_bench.py            # B200A_LIB=<other build> python tools/generic_bench.py for an A/B
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import audio_b200.transforms as T  # noqa: E402
from tools.bench_configs import time_gpu  # noqa: E402


def main():
    tag = os.path.basename(os.environ.get("B200A_LIB", "in-tree"))
    x = torch.randn(256, 160000, device="cuda")
    cases = [
        ("MelSpectrogram defaults (n_fft 400, hop 200, 128 mels) 256x10s", T.MelSpectrogram(16000).cuda(), x),
        ("MFCC defaults (n_fft 400, 40 of 128) 256x10s", T.MFCC(16000).cuda(), x),
        ("Spectrogram n_fft 1024 power=None (complex) 64x10s", T.Spectrogram(n_fft=1024, hop_length=256, power=None).cuda(), x[:64]),
    ]
    for name, mod, inp in cases:
        mod(inp)
        torch.cuda.synchronize()
        print(f"[{tag}] {name:62s} {time_gpu(lambda: mod(inp), iters=5, blocks=3):9.4f} ms", flush=True)
    spec = T.Spectrogram(n_fft=1024, hop_length=256, power=None).cuda()(x[:64])
    inv = T.InverseSpectrogram(n_fft=1024, hop_length=256).cuda()
    inv(spec, 160000)
    torch.cuda.synchronize()
    print(f"[{tag}] {'InverseSpectrogram n_fft 1024 64x10s':62s} {time_gpu(lambda: inv(spec, 160000), iters=5, blocks=3):9.4f} ms", flush=True)
    mag = spec[:8].abs().pow(2).contiguous()
    gl = T.GriffinLim(n_fft=1024, hop_length=256, length=160000, rand_init=False).cuda()
    gl(mag)
    torch.cuda.synchronize()
    print(f"[{tag}] {'GriffinLim 32 iterations n_fft 1024 8x10s':62s} {time_gpu(lambda: gl(mag), iters=2, blocks=2):9.4f} ms", flush=True)
    ps = T.PitchShift(16000, 12).cuda()
    ps(x[:16])
    torch.cuda.synchronize()
    print(f"[{tag}] {'PitchShift +12 (n_fft 512) 16x10s':62s} {time_gpu(lambda: ps(x[:16]), iters=3, blocks=2):9.4f} ms", flush=True)


if __name__ == "__main__":
    main()
