"""Pin the CPU oracle (oracle/frontend_oracle.py) against
(a) the librosa golden vectors of the reference's own tests and
(b) outputs of the reference itself (tests/golden/ref_cases.npz, made by make_golden.py).
Tolerances for (a) are the reference tests' own; for (b) the reference is an fp32
computation and the oracle fp64, so the bound is the fp32 round-off of the reference.
"""
import math

import numpy as np
import pytest
from conftest import assert_close, scaled_tol_close
from golden_cases import MEL_FB, MELSPECTROGRAM, MFCC, RESAMPLE, SPEC_VARIANTS, SPECTROGRAM

from oracle import frontend_oracle as O


def hamming(n):
    k = np.arange(n, dtype=np.float64)
    return 0.54 - 0.46 * np.cos(2 * np.pi * k / n)


def run_spec(x, n_fft=400, win_length=None, hop_length=None, pad=0, power=2.0, normalized=False, center=True,
             pad_mode="reflect", onesided=True, window=None):
    win_length = n_fft if win_length is None else win_length
    hop_length = win_length // 2 if hop_length is None else hop_length
    w = hamming(win_length) if window == "hamming" else O.hann_window(win_length)
    return O.spectrogram(x, pad, w, n_fft, hop_length, win_length, power, normalized, center, pad_mode, onesided)


# ---------------- (a) librosa goldens held by the reference's tests ---------------------------
@pytest.mark.parametrize("i", range(len(SPECTROGRAM)))
def test_spectrogram_librosa(librosa_transforms, i):
    cfg = SPECTROGRAM[i]
    got = run_spec(librosa_transforms["whitenoise"], **cfg)[0]
    assert_close(got, librosa_transforms[f"spectrogram_{i}"], rtol=1e-4, atol=1e-4, what=f"Spectrogram_{i}")


def test_spectrogram_complex_librosa(librosa_transforms):
    got = np.abs(run_spec(librosa_transforms["whitenoise"], n_fft=400, hop_length=200, power=None)[0])
    assert_close(got, librosa_transforms["spectrogram_complex"], rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("i", range(len(MELSPECTROGRAM)))
def test_melspectrogram_librosa(librosa_transforms, i):
    cfg = MELSPECTROGRAM[i]
    got = O.mel_spectrogram(librosa_transforms["sinusoid"], sample_rate=16000, **cfg)[0]
    assert_close(got, librosa_transforms[f"melspectrogram_{i:02d}"], rtol=1e-5, atol=5e-4, what=f"Mel_{i:02d}")


@pytest.mark.parametrize("i", range(len(MFCC)))
def test_mfcc_librosa(librosa_transforms, i):
    cfg = dict(MFCC[i])
    n_mfcc = cfg.pop("n_mfcc")
    got = O.mfcc(librosa_transforms["whitenoise"], 16000, n_mfcc, "ortho", False, cfg)[0]
    assert_close(got, librosa_transforms[f"mfcc_{i}"], rtol=1e-5, atol=5e-4, what=f"mfcc_{i}")


def test_power_and_magnitude_to_db_librosa(librosa_transforms):
    # get_spectrogram(n_fft=400, power=2) of data_utils.py:121-159 defaults hop to n_fft // 4
    spec = run_spec(librosa_transforms["whitenoise"], n_fft=400, hop_length=100, power=2.0)
    got = O.amplitude_to_db(spec, 10.0, 1e-10, 0.0, 80.0)[0]
    assert_close(got, librosa_transforms["power_to_db"], rtol=1e-3, atol=1e-3)
    got = O.amplitude_to_db(spec, 20.0, 1e-10, 0.0, 80.0)[0]
    assert_close(got, librosa_transforms["magnitude_to_db"], rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("i", range(len(MEL_FB)))
def test_mel_fb_librosa(librosa_melfb, i):
    c = MEL_FB[i]
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = O.melscale_fbanks(c["n_fft"] // 2 + 1, c["fmin"], c["fmax"], c["n_mels"], c["sample_rate"], c["norm"], c["mel_scale"])
    assert_close(got, librosa_melfb[f"fb_{i:02d}"], rtol=1.3e-6, atol=7e-5, what=f"mel_fb_{i:02d}")


# ---------------- (b) the reference's own outputs ----------------------------------------------
def test_config1_spectrogram(ref_cases):
    got = run_spec(ref_cases["c1_in"], n_fft=512, hop_length=256)
    assert got.shape == (1, 257, 63)
    scaled_tol_close(got, ref_cases["c1_out"], rel=2e-5, what="config 1")


@pytest.mark.parametrize("name", sorted(SPEC_VARIANTS))
def test_spectrogram_variants(ref_cases, name):
    got = run_spec(ref_cases["spec_in"], **SPEC_VARIANTS[name])
    scaled_tol_close(got, ref_cases[f"spec_{name}"], rel=3e-5, what=name)


@pytest.mark.parametrize("n_fft,hop,key", [(400, 200, "spec_complex400"), (1024, 256, "spec_complex1024")])
def test_spectrogram_complex(ref_cases, n_fft, hop, key):
    got = run_spec(ref_cases["spec_in"], n_fft=n_fft, hop_length=hop, power=None)
    ref = ref_cases[key][..., 0] + 1j * ref_cases[key][..., 1]
    scaled_tol_close(got, ref, rel=2e-5, what=key)


MEL_CASES = {
    "mel_c2_out": dict(sample_rate=16000, n_fft=1024, hop_length=256, n_mels=80),
    "mel_default_out": dict(sample_rate=16000),
    "mel_rnnt_out": dict(sample_rate=16000, n_fft=400, hop_length=160, n_mels=80),
    "mel_slaney2048_out": dict(sample_rate=22050, n_fft=2048, hop_length=512, n_mels=128, norm="slaney", mel_scale="slaney", f_max=8000.0),
    "mel_512_p1_out": dict(sample_rate=16000, n_fft=512, hop_length=128, n_mels=40, power=1.0),
    "mel_256_out": dict(sample_rate=16000, n_fft=256, hop_length=64, n_mels=80),
}


@pytest.mark.parametrize("key", sorted(MEL_CASES))
def test_melspectrogram_reference(ref_cases, key):
    got = O.mel_spectrogram(ref_cases["mel_in"], **MEL_CASES[key])
    scaled_tol_close(got, ref_cases[key], rel=3e-5, what=key)


def test_melspectrogram_scaled_rows(ref_cases):
    got = O.mel_spectrogram(ref_cases["mel_scaled_in"], sample_rate=16000, n_fft=1024, hop_length=256, n_mels=80)
    ref = ref_cases["mel_scaled_out"]
    for r in range(4):  # rows differ by 12 orders of magnitude: compare per row
        scaled_tol_close(got[r], ref[r], rel=3e-5, what=f"row {r}")
    assert np.all(got[2] == 0)


def test_mfcc_reference(ref_cases):
    kw = dict(n_fft=1024, hop_length=256, n_mels=80)
    x, xs = ref_cases["mel_in"], ref_cases["mel_scaled_in"]
    # dB features span ~[-100, 100]; DCT sums 80 of them: absolute tolerance on that scale
    assert_close(O.mfcc(x, 16000, 40, "ortho", False, kw), ref_cases["mfcc_x_out"], rtol=1e-4, atol=2e-3)
    assert_close(O.mfcc(xs, 16000, 40, "ortho", False, kw), ref_cases["mfcc_2d_out"], rtol=1e-4, atol=2e-3)
    assert_close(O.mfcc(xs[:, None, :], 16000, 40, "ortho", False, kw), ref_cases["mfcc_3d_out"], rtol=1e-4, atol=2e-3)
    assert_close(O.mfcc(xs[0], 16000, 40, "ortho", False, kw), ref_cases["mfcc_1d_out"], rtol=1e-4, atol=2e-3)
    # the batch-coupled clamp really differs from the per-item one
    assert np.abs(ref_cases["mfcc_2d_out"] - ref_cases["mfcc_3d_out"][:, 0]).max() > 1.0
    got = O.mfcc(x, 16000, 13, "ortho", True, dict(n_fft=400, hop_length=160, n_mels=23))
    assert_close(got, ref_cases["mfcc_log_out"], rtol=1e-4, atol=1e-3)
    got = O.mfcc(x, 16000, 20, None, False, dict(n_fft=512, hop_length=256, n_mels=64))
    assert_close(got, ref_cases["mfcc_nonorm_out"], rtol=1e-4, atol=2e-2)
    assert_close(O.mfcc(x), ref_cases["mfcc_default_out"], rtol=1e-4, atol=2e-3)


def test_amplitude_to_db_reference(ref_cases):
    p = ref_cases["db_in"]
    assert_close(O.amplitude_to_db(p, 10.0, 1e-10, 0.0, 80.0), ref_cases["db_power_top80_3d"], rtol=1e-5, atol=1e-4)
    assert_close(O.amplitude_to_db(p[:, None], 10.0, 1e-10, 0.0, 80.0), ref_cases["db_power_top80_4d"], rtol=1e-5, atol=1e-4)
    assert_close(O.amplitude_to_db(p, 20.0, 1e-10, 0.0, None), ref_cases["db_mag_none"], rtol=1e-5, atol=1e-4)


def test_constants_reference(ref_cases):
    fb = O.melscale_fbanks(513, 0.0, 8000.0, 80, 16000)
    # the reference builds fb in fp32; its own test allows atol=7e-5 rtol=1.3e-6 against librosa (fp64)
    assert_close(fb, ref_cases["mel_c2_fb"], rtol=1.3e-6, atol=7e-5)
    fb = O.melscale_fbanks(1025, 0.0, 8000.0, 128, 22050, "slaney", "slaney")
    assert_close(fb, ref_cases["mel_slaney2048_fb"], rtol=1.3e-6, atol=7e-5)
    # fp32 cos of arguments up to ~122 rad in the reference: ~7e-6 argument round-off
    assert_close(O.create_dct(40, 80, "ortho"), ref_cases["mfcc_dct"], rtol=1e-6, atol=5e-6)
    assert_close(O.hann_window(1024), ref_cases["mel_c2_window"], rtol=1e-6, atol=1e-7)
    k, w = O.sinc_resample_kernel(44100, 16000, 100, resampling_method="sinc_interp_kaiser")
    assert w == 17 and k.shape == (160, 475)
    assert_close(k, ref_cases["rs_kaiser_kernel"][:, 0], rtol=1e-4, atol=4e-6)  # float32 phase term of the cached kernel (functional.py:1378)
    k, w = O.sinc_resample_kernel(44100, 16000, 100)
    assert_close(k, ref_cases["rs_hann_kernel"][:, 0], rtol=1e-4, atol=4e-6)  # float32 phase term of the cached kernel (functional.py:1378)


@pytest.mark.parametrize("key", sorted(RESAMPLE))
def test_resample_reference(ref_cases, key):
    cut, kw = RESAMPLE[key]
    x = ref_cases["rs_in"] if cut is None else ref_cases["rs_in"][:, :cut]
    got = O.resample(x, **kw)
    ref = ref_cases[key]
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= 1e-4 * np.abs(ref).max(), key


def test_resample_functional_reference(ref_cases):
    x = ref_cases["rs_in"]
    got = O.resample(x, 44100, 16000, resampling_method="sinc_interp_kaiser")
    assert np.abs(got - ref_cases["rs_func_kaiser"]).max() <= 1e-4 * np.abs(got).max()
    got = O.resample(x, 3, 2)
    assert got.shape == ref_cases["rs_func_hann_3_2"].shape
    assert np.abs(got - ref_cases["rs_func_hann_3_2"]).max() <= 1e-4 * np.abs(got).max()


def test_integer_bookkeeping(ref_integers):
    for L, n_fft, hop, center, pad, t in ref_integers["stft_frames"]:
        if t < 0:
            continue  # torch refused (too short): error paths are covered in test_bookkeeping.py
        assert O.num_frames(int(L), int(n_fft), int(hop), bool(center), int(pad)) == t
    for o, n, L, w, taps, out_len in ref_integers["resample"]:
        g = math.gcd(int(o), int(n))
        o_r, n_r = int(o) // g, int(n) // g
        assert math.ceil(6 * o_r / (min(o_r, n_r) * 0.99)) == w
        assert 2 * w + o_r == taps
        assert O.resample_len(int(L), o_r, n_r) == out_len


def test_mfcc_tolerance_is_pinned(ref_cases):
    """The GPU MFCC tests use |a-e| <= 1e-4 |e| + 1e-4 rms(e).  That is the stated 1e-4 bar, not slack picked to make
    a kernel pass: the reference's OWN float32 CPU outputs (the fixtures) are compared with the float64 oracle
    under the same rule and must use at most a tenth of it (measured: 1 % - 6 %)."""
    x, xs = ref_cases["mel_in"], ref_cases["mel_scaled_in"]
    kw = dict(n_fft=1024, hop_length=256, n_mels=80)
    cases = {
        "mfcc_x_out": O.mfcc(x, 16000, 40, "ortho", False, kw),
        "mfcc_2d_out": O.mfcc(xs, 16000, 40, "ortho", False, kw),
        "mfcc_3d_out": O.mfcc(xs[:, None, :], 16000, 40, "ortho", False, kw),
        "mfcc_log_out": O.mfcc(x, 16000, 13, "ortho", True, dict(n_fft=400, hop_length=160, n_mels=23)),
        "mfcc_nonorm_out": O.mfcc(x, 16000, 20, None, False, dict(n_fft=512, hop_length=256, n_mels=64)),
        "mfcc_default_out": O.mfcc(x),
    }
    for key, exp in cases.items():
        err = np.abs(ref_cases[key] - exp)
        tol = 1e-4 * np.abs(exp) + 1e-4 * np.sqrt(np.mean(exp**2))
        worst = float((err / tol).max())
        assert worst < 0.1, f"{key}: the reference's fp32 run uses {worst:.2f} of the 1e-4 rule"
