#!/usr/bin/env python
"""Generate tests/golden/istft_ref_cases.npz: outputs of the reference's F.inverse_spectrogram (over torch.istft, CPU,
float32) on generic complex spectrograms (a forward STFT of noise, perturbed so that it is NOT a consistent STFT),
with the inputs stored next to them.  Needs /root/reference; run once in the build container:

    python tests/golden/make_istft_golden.py
"""
import json
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(REF, "src"))
import torchaudio  # noqa: E402
import torchaudio.functional as F  # noqa: E402

assert torchaudio.__file__.startswith(REF), torchaudio.__file__

CASES = [
    dict(n_fft=512, hop=128, win=512, normalized=False, center=True, length=None, pad=0),
    dict(n_fft=400, hop=200, win=400, normalized=True, center=True, length=7900, pad=0),
    dict(n_fft=1024, hop=256, win=800, normalized="frame_length", center=True, length=7500, pad=30),
    dict(n_fft=256, hop=64, win=256, normalized=False, center=False, length=None, pad=0),
    dict(n_fft=100, hop=25, win=100, normalized="window", center=True, length=5100, pad=0),
    dict(n_fft=512, hop=256, win=512, normalized=False, center=True, length=8400, pad=0),  # longer than covered: zero tail
]


def main():
    g = torch.Generator().manual_seed(77)
    out = {"cases": np.array([json.dumps(c) for c in CASES])}
    for i, c in enumerate(CASES):
        x = torch.randn(3, 8000, generator=g)
        window = torch.hann_window(c["win"]) if c["center"] else torch.hamming_window(c["win"])
        spec = F.spectrogram(x, c["pad"], window, c["n_fft"], c["hop"], c["win"], None, c["normalized"], c["center"])
        spec = spec * (1 + 0.1 * torch.randn(spec.shape, generator=g))
        y = F.inverse_spectrogram(spec, c["length"], c["pad"], window, c["n_fft"], c["hop"], c["win"], c["normalized"],
                                  c["center"])
        out[f"spec_{i}"], out[f"window_{i}"], out[f"out_{i}"] = spec.numpy(), window.numpy(), y.numpy()
        print(i, tuple(spec.shape), "->", tuple(y.shape))
    np.savez_compressed(os.path.join(HERE, "istft_ref_cases.npz"), **out)


if __name__ == "__main__":
    main()
