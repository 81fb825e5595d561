"""Drop-in ``torchaudio.functional`` surface of the hot path, backed by libb200audio.so.

Same names, argument order, defaults and error behaviour as the reference
(/root/reference/src/torchaudio/functional/functional.py):
``spectrogram`` (54-145), ``melscale_fbanks`` (518-587), ``linear_fbanks`` (590-633),
``create_dct`` (636-667), ``amplitude_to_DB`` (356-404), ``resample`` (1435-1490) and the two
private helpers ``transforms`` imports (``_get_sinc_resample_kernel`` 1305-1402,
``_apply_sinc_resample_kernel`` 1405-1432).

Differences, all explicit (never a silent fallback): CUDA float32 tensors only, forward only.
"""
from __future__ import annotations

import collections
import math
import warnings
from typing import Optional, Union

import torch
from torch import Tensor

from . import _lib
from ._bookkeeping import resample_ratio
from ._constants import create_dct, linear_fbanks, melscale_fbanks, sinc_resample_kernel
from ._plans import FrontendPlan, ResamplePlan, _no_autograd, _require_cuda_f32, _stream_ptr, new_group_max

__all__ = [
    "spectrogram",
    "inverse_spectrogram",
    "griffinlim",
    "phase_vocoder",
    "pitch_shift",
    "melscale_fbanks",
    "linear_fbanks",
    "create_dct",
    "amplitude_to_DB",
    "resample",
    "speed",
    "spectral_centroid",
    "mel_spectrogram",
    "mfcc",
]


_PLAN_CACHE: "dict[tuple, FrontendPlan]" = {}


def _plan_for(desc, device) -> FrontendPlan:
    """One FrontendPlan per (descriptor, device) for the functional entry points, so that repeated calls reuse the
    prepared workspace (the plan itself re-prepares when the window / filterbank tensors change)."""
    key = (tuple(None if (isinstance(v, float) and v != v) else v for v in desc.key()), str(device))  # NaN power -> None
    plan = _PLAN_CACHE.get(key)
    if plan is None:
        if len(_PLAN_CACHE) >= 64:
            _PLAN_CACHE.pop(next(iter(_PLAN_CACHE)))
        plan = _PLAN_CACHE[key] = FrontendPlan(desc)
    return plan


def _get_spec_norms(normalized: Union[str, bool]):
    """(frame_length_norm, window_norm) -- reference functional.py:228-242."""
    if isinstance(normalized, str):
        if normalized not in ("frame_length", "window"):
            raise ValueError("Invalid normalized parameter: {}".format(normalized))
        return normalized == "frame_length", normalized == "window"
    if isinstance(normalized, bool):
        return False, normalized
    raise TypeError("Input type not supported")


def _unpack(out: Tensor, waveform: Tensor) -> Tensor:
    """(rows, T, W[,2]) frame-major -> logical (..., W, T) view, as the reference returns it."""
    lead = waveform.shape[:-1]
    if out.dim() == 4:  # complex
        out = torch.view_as_complex(out)
    return out.reshape(lead + out.shape[-2:]).transpose(-1, -2)


def spectrogram(
    waveform: Tensor,
    pad: int,
    window: Tensor,
    n_fft: int,
    hop_length: int,
    win_length: int,
    power: Optional[float],
    normalized: Union[bool, str],
    center: bool = True,
    pad_mode: str = "reflect",
    onesided: bool = True,
    return_complex: Optional[bool] = None,
) -> Tensor:
    """``(..., time) -> (..., freq, time)``; one fused kernel (pad, frame, window, FFT, |.|^p)."""
    if return_complex is not None:
        warnings.warn(
            "`return_complex` argument is now deprecated and is not effective."
            "`torchaudio.functional.spectrogram(power=None)` always returns a tensor with "
            "complex dtype. Please remove the argument in the function call."
        )
    fl_norm, win_norm = _get_spec_norms(normalized)
    desc = FrontendPlan.make_desc(n_fft, win_length, hop_length, pad, center, pad_mode, onesided, fl_norm, win_norm, power)
    plan = _plan_for(desc, waveform.device)
    ws = plan.workspace(window, None, None)
    stage = _lib.STAGE_COMPLEX if power is None else _lib.STAGE_POWER
    return _unpack(plan.run(ws, stage, waveform), waveform)


# ---- inverse spectrogram ------------------------------------------------------------------------------------
_ENVELOPE_OK: "collections.OrderedDict[tuple, Tensor]" = collections.OrderedDict()
_ENVELOPE_CACHE_SIZE = 64


def _check_window_envelope(window: Tensor, n_fft: int, win_length: int, hop: int, frames: int, start: int, end: int) -> None:
    """torch.istft refuses windows whose overlap-added square dips below 1e-11 inside the returned range
    ("window overlap add min"); it finds out with a device synchronisation, and so does this check -- once per
    (window tensor, geometry).  The cache entry HOLDS the window tensor, so its (data_ptr, _version) key cannot
    be matched by a different window that was handed the recycled allocation; the cache is a bounded LRU."""
    key = (window.data_ptr(), -1 if window.is_inference() else window._version, str(window.device), n_fft, win_length,
           hop, frames, start, end)
    if key in _ENVELOPE_OK:
        _ENVELOPE_OK.move_to_end(key)
        return
    import numpy as np

    w = np.zeros(n_fft, dtype=np.float64)
    left = (n_fft - win_length) // 2
    w[left:left + win_length] = window.detach().double().cpu().numpy()
    expected = n_fft + hop * (frames - 1)
    # overlap-added w^2 without a Python loop over frames: scatter-add over the (frames, n_fft) index grid
    env = np.zeros(expected)
    idx = (np.arange(frames)[:, None] * hop + np.arange(n_fft)[None, :]).ravel()
    np.add.at(env, idx, np.tile(w * w, frames))
    seg = env[start:min(end, expected)]
    if seg.size and np.abs(seg).min() < 1e-11:
        raise RuntimeError("istft(...) window overlap add min: 1 (the window envelope is zero inside the output range)")
    _ENVELOPE_OK[key] = window
    if len(_ENVELOPE_OK) > _ENVELOPE_CACHE_SIZE:
        _ENVELOPE_OK.popitem(last=False)


def inverse_spectrogram(
    spectrogram: Tensor,
    length: Optional[int],
    pad: int,
    window: Tensor,
    n_fft: int,
    hop_length: int,
    win_length: int,
    normalized: Union[bool, str],
    center: bool = True,
    pad_mode: str = "reflect",
    onesided: bool = True,
) -> Tensor:
    """``(..., freq, time)`` complex64 -> ``(..., time)``: least-squares inverse of ``spectrogram(power=None)``
    (reference functional.py:148-225 over ``torch.istft``).  Two kernels: Hermitian inverse FFT x window per frame pair,
    then overlap-add with window-envelope normalisation."""
    fl_norm, win_norm = _get_spec_norms(normalized)
    if not spectrogram.is_complex():
        raise ValueError("Expected `spectrogram` to be complex dtype.")
    if not spectrogram.is_cuda:
        raise RuntimeError(
            f"audio_b200: spectrogram is on '{spectrogram.device}'. This package runs only hand-written sm_100a CUDA "
            "kernels; there is no CPU or ATen fallback -- move the tensor (and the module) to a CUDA device."
        )
    if spectrogram.dtype != torch.complex64:
        raise TypeError(f"audio_b200: spectrogram must be complex64 (got {spectrogram.dtype})")
    if not onesided:
        raise NotImplementedError("audio_b200: inverse_spectrogram(onesided=False) is not implemented")
    _no_autograd(spectrogram)
    shape = spectrogram.size()
    n_bins, frames = shape[-2], shape[-1]
    if n_bins != n_fft // 2 + 1:
        raise RuntimeError(f"istft: expected {n_fft // 2 + 1} frequency bins for n_fft={n_fft}, got {n_bins}")
    spec3 = spectrogram.reshape(-1, n_bins, frames)
    rows = spec3.shape[0]
    desc = FrontendPlan.make_desc(n_fft, win_length, hop_length, 0, center, "reflect", True, fl_norm, win_norm, 2.0)
    plan = _plan_for(desc, spectrogram.device)
    ws = plan.workspace(window, None, None)
    expected = n_fft + hop_length * (frames - 1)
    start = n_fft // 2 if center else 0
    if length is not None:
        out_len = length + 2 * pad
    else:
        out_len = expected - 2 * start if center else expected
    if out_len <= 0:
        raise RuntimeError(f"istft: the requested signal is empty (frames={frames}, n_fft={n_fft})")
    _check_window_envelope(window, n_fft, win_length, hop_length, frames, start, start + out_len)
    if start + out_len > expected:
        warnings.warn("The length of signal is shorter than the length parameter. Result is being padded with zeros in "
                      "the tail. Please check your center and hop_length settings.")
    dev = spectrogram.device
    real = torch.view_as_real(spec3)  # (rows, bins, frames, 2) float32 view, same storage
    with torch.cuda.device(dev):
        frame_buf = torch.empty((rows, frames, n_fft), dtype=torch.float32, device=dev)
        out = torch.empty((rows, out_len), dtype=torch.float32, device=dev)
        rc = _lib.lib().b200a_istft_run(
            desc, ws.data_ptr(), real.data_ptr(), rows, frames, spec3.stride(0), spec3.stride(1), spec3.stride(2),
            frame_buf.data_ptr(), out.data_ptr(), out_len, start, out_len, _stream_ptr(dev),
        )
    _lib.check(rc, "istft_run")
    if length is not None and pad > 0:
        out = out[:, pad:-pad]
    return out.reshape(shape[:-2] + out.shape[-1:])


def griffinlim(
    specgram: Tensor,
    window: Tensor,
    n_fft: int,
    hop_length: int,
    win_length: int,
    power: float,
    n_iter: int,
    momentum: float,
    length: Optional[int],
    rand_init: bool,
) -> Tensor:
    """Fast Griffin-Lim phase recovery: ``(..., freq, time)`` |X|^power -> ``(..., time)`` (reference
    functional.py:255-353).  Every iteration is four launches on device-resident buffers: the phase step
    (``b200a_griffinlim_update``), the two inverse-STFT kernels, and the fused forward STFT (complex stage)."""
    if not 0 <= momentum < 1:
        raise ValueError("momentum must be in range [0, 1). Found: {}".format(momentum))
    momentum = momentum / (1 + momentum)
    _require_cuda_f32(specgram, "specgram")
    _no_autograd(specgram)
    shape = specgram.size()
    n_bins, frames = shape[-2], shape[-1]
    spec3 = specgram.reshape(-1, n_bins, frames)
    rows = spec3.shape[0]
    dev = specgram.device
    desc = FrontendPlan.make_desc(n_fft, win_length, hop_length, 0, True, "reflect", True, False, False, None)
    plan = _plan_for(desc, dev)
    ws = plan.workspace(window, None, None)
    expected = n_fft + hop_length * (frames - 1)
    start = n_fft // 2
    out_len = length if length is not None else expected - 2 * start
    _check_window_envelope(window, n_fft, win_length, hop_length, frames, start, start + out_len)
    lib = _lib.lib()
    stream = _stream_ptr(dev)
    with torch.cuda.device(dev):
        proj = torch.empty((rows, frames, n_bins, 2), dtype=torch.float32, device=dev)
        frame_buf = torch.empty((rows, frames, n_fft), dtype=torch.float32, device=dev)
        wave = torch.empty((rows, out_len), dtype=torch.float32, device=dev)
        rebuilt, tprev = None, None
        first_raw = False
        if rand_init:
            # the reference's own call (functional.py:310-311): uniform real and imaginary parts from torch's generator
            # on this device, in the (rows, freq, time) element order; used un-normalised for the first inversion
            init = torch.rand(spec3.size(), dtype=torch.complex64, device=dev)
            rebuilt = torch.view_as_real(init.transpose(1, 2).contiguous())
            first_raw = True

        def invert():
            rc = lib.b200a_istft_run(desc, ws.data_ptr(), proj.data_ptr(), rows, frames, frames * n_bins, 1, n_bins,
                                     frame_buf.data_ptr(), wave.data_ptr(), out_len, start, out_len, stream)
            _lib.check(rc, "istft_run")

        def step(raw=False):
            rc = lib.b200a_griffinlim_update(
                spec3.data_ptr(), spec3.stride(0), spec3.stride(1), spec3.stride(2), 1.0 / float(power),
                None if rebuilt is None else rebuilt.data_ptr(), None if tprev is None or not momentum else tprev.data_ptr(),
                float(momentum), 0 if raw else 1, proj.data_ptr(), rows, n_bins, frames, stream)
            _lib.check(rc, "griffinlim_update")

        for it in range(n_iter):
            step(raw=first_raw and it == 0)
            invert()
            new = plan.run(ws, _lib.STAGE_COMPLEX, wave)  # (rows, T', bins, 2) frame-major
            if new.shape[1] != frames:
                raise RuntimeError(
                    f"griffinlim: the rebuilt spectrogram has {new.shape[1]} frames, the input {frames} "
                    "(`length` is inconsistent with hop_length and the number of frames)"
                )
            tprev, rebuilt = (None if (first_raw and it == 0) else rebuilt), new
        step(raw=first_raw and n_iter == 0)
        invert()
    return wave.reshape(shape[:-2] + wave.shape[-1:])


def phase_vocoder(complex_specgrams: Tensor, rate: float, phase_advance: Tensor) -> Tensor:
    """Stretch a complex spectrogram in time by ``rate`` without modifying pitch: ``(..., freq, num_frame)`` ->
    ``(..., freq, ceil(num_frame / rate))`` (reference functional.py:713-803)."""
    if rate == 1.0:
        return complex_specgrams
    if not complex_specgrams.is_complex():
        raise ValueError("audio_b200: phase_vocoder expects a complex spectrogram")
    if not complex_specgrams.is_cuda:
        raise RuntimeError(
            f"audio_b200: complex_specgrams is on '{complex_specgrams.device}'. This package runs only hand-written "
            "sm_100a CUDA kernels; there is no CPU or ATen fallback -- move the tensor (and the module) to a CUDA device."
        )
    if complex_specgrams.dtype != torch.complex64:
        raise TypeError(f"audio_b200: complex_specgrams must be complex64 (got {complex_specgrams.dtype})")
    _no_autograd(complex_specgrams)
    _require_cuda_f32(phase_advance, "phase_advance")
    shape = complex_specgrams.size()
    n_bins, frames = shape[-2], shape[-1]
    spec3 = complex_specgrams.reshape(-1, n_bins, frames)
    rows = spec3.shape[0]
    pa = phase_advance.reshape(-1).contiguous()
    if pa.numel() != n_bins:
        raise RuntimeError(f"phase_advance must have one entry per frequency bin ({n_bins}), got {pa.numel()}")
    frames_out = int(math.ceil(frames / rate))  # len(torch.arange(0, frames, rate))
    dev = complex_specgrams.device
    with torch.cuda.device(dev):
        out = torch.empty((rows, frames_out, n_bins, 2), dtype=torch.float32, device=dev)
        rc = _lib.lib().b200a_phase_vocoder(
            torch.view_as_real(spec3).data_ptr(), spec3.stride(0), spec3.stride(1), spec3.stride(2), rows, n_bins, frames,
            float(rate), pa.data_ptr(), out.data_ptr(), frames_out, _stream_ptr(dev))
    _lib.check(rc, "phase_vocoder")
    res = torch.view_as_complex(out).transpose(-1, -2)  # logical (rows, freq, frames_out) over the frame-major buffer
    return res.reshape(shape[:-2] + res.shape[1:])


def pitch_shift(
    waveform: Tensor,
    sample_rate: int,
    n_steps: int,
    bins_per_octave: int = 12,
    n_fft: int = 512,
    win_length: Optional[int] = None,
    hop_length: Optional[int] = None,
    window: Optional[Tensor] = None,
) -> Tensor:
    """Shift the pitch of a waveform by ``n_steps`` steps (reference functional.py:1579-1719): STFT -> phase vocoder
    (rate 2^(-n_steps / bins_per_octave)) -> inverse STFT -> resample back to the original duration -> crop / zero-pad
    to the input length.  Five kernels of this library, no host round trip."""
    _require_cuda_f32(waveform, "waveform")
    if hop_length is None:
        hop_length = n_fft // 4
    if win_length is None:
        win_length = n_fft
    if window is None:
        window = torch.hann_window(window_length=win_length, device=waveform.device)
    shape = waveform.size()
    flat = waveform.reshape(-1, shape[-1])
    ori_len = shape[-1]
    rate = 2.0 ** (-float(n_steps) / bins_per_octave)
    spec_f = spectrogram(flat, 0, window, n_fft, hop_length, win_length, None, False)
    phase_advance = torch.linspace(0, math.pi * hop_length, spec_f.shape[-2], device=spec_f.device)[..., None]
    spec_stretch = phase_vocoder(spec_f, rate, phase_advance)
    len_stretch = int(round(ori_len / rate))
    stretched = inverse_spectrogram(spec_stretch, len_stretch, 0, window, n_fft, hop_length, win_length, False)
    shifted = resample(stretched, int(sample_rate / rate), sample_rate)
    shift_len = shifted.size()[-1]
    if shift_len > ori_len:
        shifted = shifted[..., :ori_len]
    else:
        shifted = torch.nn.functional.pad(shifted, [0, ori_len - shift_len])
    return shifted.reshape(shape[:-1] + shifted.shape[-1:])


def _db_groups(shape) -> int:
    """How many independent top_db cut-offs the reference uses for a tensor of this shape
    (functional.py:395-399: dims beyond the last three are separate items)."""
    groups = 1
    for s in shape[:-3]:
        groups *= s
    return groups


def amplitude_to_DB(
    x: Tensor, multiplier: float, amin: float, db_multiplier: float, top_db: Optional[float] = None
) -> Tensor:
    _require_cuda_f32(x, "x")
    _no_autograd(x)
    xc = x.contiguous()
    out = torch.empty_like(xc)
    if xc.numel() == 0:
        return out
    groups = _db_groups(xc.shape) if top_db is not None else 1
    dev = xc.device
    with torch.cuda.device(dev):
        scratch = torch.empty(groups, dtype=torch.float32, device=dev)
        rc = _lib.lib().b200a_amplitude_to_db(
            xc.data_ptr(), groups, xc.numel() // groups, float(multiplier), float(amin),
            float(multiplier) * float(db_multiplier), -1.0 if top_db is None else float(top_db),
            scratch.data_ptr(), out.data_ptr(), _stream_ptr(dev),
        )
    _lib.check(rc, "amplitude_to_db")
    return out


def _apply_fbank(specgram: Tensor, fb: Tensor) -> Tensor:
    """MelScale.forward on an existing spectrogram of logical shape (..., n_bins, T)."""
    _require_cuda_f32(specgram, "specgram")
    _require_cuda_f32(fb, "fb")
    _no_autograd(specgram)
    n_bins, frames = specgram.shape[-2], specgram.shape[-1]
    if fb.shape[0] != n_bins:
        raise RuntimeError(f"mat1 and mat2 shapes cannot be multiplied: n_bins={n_bins} vs fb {tuple(fb.shape)}")
    lead = specgram.shape[:-2]
    s3 = specgram.reshape((-1, n_bins, frames))
    rows = s3.shape[0]
    fbc = fb.contiguous()
    dev = specgram.device
    with torch.cuda.device(dev):
        out = torch.empty((rows, frames, fb.shape[1]), dtype=torch.float32, device=dev)
        rc = _lib.lib().b200a_apply_fbank(
            s3.data_ptr(), rows, n_bins, frames, s3.stride(0), s3.stride(1), s3.stride(2),
            fbc.data_ptr(), fb.shape[1], out.data_ptr(), _stream_ptr(dev),
        )
    _lib.check(rc, "apply_fbank")
    return out.reshape(lead + out.shape[-2:]).transpose(-1, -2)


# ---- fused MelSpectrogram / MFCC (what the transforms call) -----------------------------------
def mel_spectrogram(plan: FrontendPlan, window: Tensor, fb: Tensor, waveform: Tensor) -> Tensor:
    """Spectrogram + MelScale in ONE kernel: ``(..., time) -> (..., n_mels, time)``."""
    ws = plan.workspace(window, fb, None)
    return _unpack(plan.run(ws, _lib.STAGE_MEL, waveform), waveform)


def mfcc(
    plan: FrontendPlan,
    window: Tensor,
    fb: Tensor,
    dct_mat: Tensor,
    waveform: Tensor,
    top_db: Optional[float],
    log_mels: bool,
    process_group=None,
) -> Tensor:
    """MelSpectrogram -> dB/log -> DCT: fused front-end kernel + clamp/DCT kernel.

    The only cross-utterance coupling on the path is AmplitudeToDB's ``top_db`` clamp
    (reference functional.py:395-399): for a waveform of dim <= 2 ONE maximum is shared by the
    whole batch, for dim >= 3 each leading item has its own.  ``process_group`` (optional)
    extends the shared maximum across ranks with one all-reduce(MAX) of that scalar.
    """
    ws = plan.workspace(window, fb, dct_mat)
    rows = 1
    for s in waveform.shape[:-1]:
        rows *= s
    clamp = (not log_mels) and top_db is not None
    if clamp:
        rows_per_group = waveform.shape[-2] if waveform.dim() >= 2 else 1
        rows_per_group = max(int(rows_per_group), 1)
        groups = max((rows + rows_per_group - 1) // rows_per_group, 1)
        gmax = new_group_max(groups, waveform.device)
    else:
        rows_per_group, gmax = 1, None
    feat = plan.run(ws, _lib.STAGE_FEAT, waveform, gmax, rows_per_group)
    if clamp and waveform.dim() <= 2:
        gmax = _exchange_group_max(gmax, process_group)
    out = plan.mfcc_finish(ws, feat, gmax, rows_per_group, top_db if clamp else None)
    return _unpack(out, waveform)


def _exchange_group_max(gmax: Tensor, process_group) -> Tensor:
    """The path's only cross-rank message: all-reduce(MAX) of the running dB maximum of a 2-D batch that is sharded
    over ``process_group`` (reference semantics: ONE ``amax`` over the whole batch, functional.py:395-399).
    In place; a ``None`` group (single process) is the identity."""
    if process_group is not None:
        import torch.distributed as dist

        dist.all_reduce(gmax, op=dist.ReduceOp.MAX, group=process_group)
    return gmax


# ---- resampling ---------------------------------------------------------------------------------
def _get_sinc_resample_kernel(
    orig_freq: int,
    new_freq: int,
    gcd: int,
    lowpass_filter_width: int = 6,
    rolloff: float = 0.99,
    resampling_method: str = "sinc_interp_hann",
    beta: Optional[float] = None,
    device: torch.device = torch.device("cpu"),
    dtype: Optional[torch.dtype] = None,
):
    return sinc_resample_kernel(
        orig_freq, new_freq, gcd, lowpass_filter_width, rolloff, resampling_method, beta, device, dtype
    )


def _apply_sinc_resample_kernel(
    waveform: Tensor, orig_freq: int, new_freq: int, gcd: int, kernel: Tensor, width: int, plan: Optional[ResamplePlan] = None
) -> Tensor:
    if not waveform.is_floating_point():
        raise TypeError(f"Expected floating point type for waveform tensor, but received {waveform.dtype}.")
    if plan is None:
        plan = ResamplePlan(int(orig_freq) // gcd, int(new_freq) // gcd, width)
    return plan.run(kernel, waveform)


def resample(
    waveform: Tensor,
    orig_freq: int,
    new_freq: int,
    lowpass_filter_width: int = 6,
    rolloff: float = 0.99,
    resampling_method: str = "sinc_interp_hann",
    beta: Optional[float] = None,
) -> Tensor:
    if orig_freq <= 0.0 or new_freq <= 0.0:
        raise ValueError("Original frequency and desired frequecy should be positive")
    if orig_freq == new_freq:
        return waveform
    if not waveform.is_floating_point():
        raise TypeError(f"Expected floating point type for waveform tensor, but received {waveform.dtype}.")
    _require_cuda_f32(waveform, "waveform")
    gcd = math.gcd(int(orig_freq), int(new_freq))
    # the reference builds the taps on the waveform's device in its dtype (functional.py:1478-1488);
    # that is a handful of tiny torch elementwise launches at call time -- table building, not the hot path
    kernel, width = sinc_resample_kernel(
        orig_freq, new_freq, gcd, lowpass_filter_width, rolloff, resampling_method, beta, waveform.device, waveform.dtype
    )
    return _apply_sinc_resample_kernel(waveform, orig_freq, new_freq, gcd, kernel, width)


def speed(waveform: Tensor, orig_freq: int, factor: float, lengths: Optional[Tensor] = None):
    """Adjusts waveform speed (reference functional.py:2384-2423): ``resample`` from ``int(factor * orig_freq)`` to
    ``orig_freq``; returns ``(waveform', lengths')``."""
    source, target = int(factor * orig_freq), int(orig_freq)
    g = math.gcd(source, target)
    source, target = source // g, target // g
    out_lengths = None if lengths is None else torch.ceil(lengths * target / source).to(lengths.dtype)
    return resample(waveform, source, target), out_lengths


# ---- spectral centroid (a weighted-sum epilogue of the same fused kernel) --------------
def spectral_centroid(
    waveform: Tensor,
    sample_rate: int,
    pad: int,
    window: Tensor,
    n_fft: int,
    hop_length: int,
    win_length: int,
) -> Tensor:
    """``(..., time) -> (..., frames)``: sum_k f_k |X_k| / sum_k |X_k| (reference functional.py:1257-1299).

    The magnitude spectrogram is contracted inside the fused kernel with the two-column matrix
    ``[bin frequency | 1]`` (the mel stage with a 2-filter bank), then one tiny kernel divides the pair.
    """
    _require_cuda_f32(waveform, "waveform")
    desc = FrontendPlan.make_desc(n_fft, win_length, hop_length, pad, True, "reflect", True, False, False, 1.0, n_mels=2)
    plan = FrontendPlan(desc)
    dev = waveform.device
    freqs = torch.linspace(0, sample_rate // 2, steps=1 + n_fft // 2, device=dev)
    fb = torch.stack([freqs, torch.ones_like(freqs)], dim=1).contiguous()
    ws = plan.workspace(window, fb, None)
    pairs = plan.run(ws, _lib.STAGE_MEL, waveform)  # (rows, T, 2)
    rows, frames, _ = pairs.shape
    with torch.cuda.device(dev):
        out = torch.empty((rows, frames), dtype=torch.float32, device=dev)
        rc = _lib.lib().b200a_ratio_f32(pairs.data_ptr(), rows * frames, out.data_ptr(), _stream_ptr(dev))
    _lib.check(rc, "ratio_f32")
    return out.reshape(waveform.shape[:-1] + (frames,))
