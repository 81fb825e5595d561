import sys
sys.path.insert(0, "/root/repo")
import torch
import audio_b200.transforms as T
import audio_b200.compliance.kaldi as K
x = torch.randn(3, 12000, device="cuda")
for n_fft in (256, 512, 1024):
    y = T.MelSpectrogram(16000, n_fft=n_fft, hop_length=n_fft // 4, n_mels=40).cuda()(x)
y = T.MFCC(16000, n_mfcc=13, melkwargs=dict(n_fft=512, hop_length=160, n_mels=40)).cuda()(x)
y = T.Spectrogram(n_fft=1024, hop_length=256).cuda()(x)
y = K.fbank_batch(x * 1000, num_mel_bins=40, snip_edges=False, use_energy=True)
y = T.Resample(44100, 16000).cuda()(x)
torch.cuda.synchronize()
print("done")
