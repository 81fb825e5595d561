"""Small invocations of every kernel family for compute-sanitizer:
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py
(racecheck reports the mbarrier-synchronised hand-offs of the warp-specialised kernels as hazards: it does not model
mbarrier / tcgen05.commit ordering; memcheck and initcheck are the meaningful tools here)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import audio_b200.compliance.kaldi as K  # noqa: E402
import audio_b200.transforms as T  # noqa: E402

x = torch.randn(3, 12000, device="cuda")
for n_fft in (256, 512, 1024, 2048, 400):
    T.MelSpectrogram(16000, n_fft=n_fft, hop_length=n_fft // 4, n_mels=40).cuda()(x)
    T.Spectrogram(n_fft=n_fft, hop_length=n_fft // 4).cuda()(x)
    spec = T.Spectrogram(n_fft=n_fft, hop_length=n_fft // 4, power=None).cuda()(x)
    T.InverseSpectrogram(n_fft=n_fft, hop_length=n_fft // 4).cuda()(spec, 12000)
T.MFCC(16000, n_mfcc=13, melkwargs=dict(n_fft=512, hop_length=160, n_mels=40)).cuda()(x)
T.MFCC(16000, n_mfcc=40, melkwargs=dict(n_fft=1024, hop_length=256, n_mels=80)).cuda()(x.reshape(1, 3, -1))
for kw in (dict(num_mel_bins=40, snip_edges=False, use_energy=True), dict(num_mel_bins=23), dict(frame_length=20.0, round_to_power_of_two=False)):
    K.fbank_batch(x * 1000, **kw)
K.mfcc_batch(x * 1000, subtract_mean=True)
K.spectrogram_batch(x * 1000)
T.Resample(44100, 16000).cuda()(x)
T.Resample(16000, 22050, resampling_method="sinc_interp_kaiser").cuda()(x)
T.GriffinLim(n_fft=512, hop_length=128, n_iter=3, length=12000, rand_init=False).cuda()(
    T.Spectrogram(n_fft=512, hop_length=128).cuda()(x))
T.PitchShift(16000, 12).cuda()(x)
T.TimeStretch(hop_length=128, n_freq=257, fixed_rate=1.3).cuda()(T.Spectrogram(n_fft=512, hop_length=128, power=None).cuda()(x))
torch.cuda.synchronize()
print("done")
