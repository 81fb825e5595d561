import sys, os
sys.path.insert(0, "/root/repo")
import torch
import audio_b200.transforms as T
from tools.bench_configs import time_gpu
x = torch.randn(256, 160000, device="cuda")
for n_fft in (1024, 512, 256):
    m = T.Spectrogram(n_fft=n_fft, hop_length=n_fft // 4).cuda()
    m(x); torch.cuda.synchronize()
    print("spec", n_fft, f"{time_gpu(lambda: m(x)):.4f} ms", flush=True)
