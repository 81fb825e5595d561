#!/usr/bin/env python
"""Mel stage of n_fft = 256 / 512 / 1024: tcgen05 contraction (default) vs the mma.sync TF32x3 path (B200A_TC=0).

The switch is read once per process, so this script runs itself twice and compares the saved outputs:
    python tools/tc_check.py            # prints max relative difference and both kernel times
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(tag):
    import numpy as np
    import torch

    import audio_b200.transforms as T
    from tools.bench_configs import time_gpu

    torch.manual_seed(0)
    res = {}
    for n_fft, n_mels, batch in ((1024, 80, 256), (1024, 128, 64), (1024, 23, 8), (512, 80, 256), (512, 40, 16),
                                 (256, 80, 256), (256, 40, 64), (256, 128, 32), (256, 23, 8)):
        x = torch.randn(batch, 160000, device="cuda")
        m = T.MelSpectrogram(16000, n_fft=n_fft, hop_length=n_fft // 4, n_mels=n_mels).cuda()
        y = m(x)
        torch.cuda.synchronize()
        res[f"mel{n_fft}_{n_mels}"] = y[:4].cpu().numpy()
        t = time_gpu(lambda: m(x))
        print(f"[{tag}] MelSpectrogram n_fft={n_fft} hop={n_fft // 4} n_mels={n_mels} batch={batch}: {t:.4f} ms", flush=True)
    x = torch.randn(8, 3, 20000, device="cuda")
    f = T.MFCC(16000, n_mfcc=13, melkwargs=dict(n_fft=256, hop_length=80, n_mels=40)).cuda()
    res["mfcc256"] = f(x).cpu().numpy()
    f = T.MFCC(16000, n_mfcc=40, melkwargs=dict(n_fft=1024, hop_length=256, n_mels=80)).cuda()
    res["mfcc1024"] = f(x).cpu().numpy()
    np.savez(os.path.join(ROOT, "gpurun_out", f"tc_{tag}.npz"), **res)


def main():
    if len(sys.argv) > 1:
        return child(sys.argv[1])
    import numpy as np

    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    for tag, val in (("tc", "1"), ("mma", "0")):
        env = dict(os.environ, B200A_TC=val)
        subprocess.run([sys.executable, os.path.abspath(__file__), tag], env=env, check=True, timeout=600)
    a = np.load(os.path.join(ROOT, "gpurun_out", "tc_tc.npz"))
    b = np.load(os.path.join(ROOT, "gpurun_out", "tc_mma.npz"))
    for k in a.files:
        scale = np.abs(b[k]).max()
        print(f"{k}: max |tc - mma| / max|mma| = {np.abs(a[k] - b[k]).max() / scale:.3e}")


if __name__ == "__main__":
    main()
