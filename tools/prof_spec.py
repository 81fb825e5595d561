import sys
sys.path.insert(0, ".")
import torch
import audio_b200.transforms as T
x = torch.randn(256, 160000, device="cuda")
m = T.Spectrogram(n_fft=1024, hop_length=256).cuda()
for _ in range(3):
    m(x)
torch.cuda.synchronize()
