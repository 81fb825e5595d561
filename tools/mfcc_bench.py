import sys, torch
sys.path.insert(0, ".")
import audio_b200.transforms as T
dev = "cuda:0"
mf = T.MFCC(16000, n_mfcc=40, melkwargs=dict(n_fft=1024, hop_length=256, n_mels=80)).to(dev)
x = torch.randn(256, 160000, device=dev)
with torch.inference_mode():
    for _ in range(3): y = mf(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): y = mf(x)
    e1.record(); torch.cuda.synchronize()
print("mfcc ms", e0.elapsed_time(e1) / 20)
