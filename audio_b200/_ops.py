"""``torch.library`` registration of the three hot-path ops, so the dispatcher, ``torch.profiler`` and CUDA-graph
capture tooling see them as ``b200audio::frontend_run`` / ``mfcc_finish`` / ``resample_run``.

Same shape as the reference's native ops -- ``STABLE_TORCH_LIBRARY_FRAGMENT(torchaudio, m){ m.def(...) }`` with a
per-backend ``..._IMPL(torchaudio, CUDA, m)`` (/root/reference/src/libtorchaudio/lfilter.cpp:118-138) bound on the
Python side as ``torch.ops.torchaudio.X`` (/root/reference/src/torchaudio/functional/filtering.py:935-938): here the
schema is defined from Python (the library itself has no torch headers), the CUDA implementation forwards to the
C ABI of libb200audio.so on the current stream, and a Meta implementation gives shapes for fake tensors.  There is
deliberately NO CPU implementation: a CPU tensor fails in the dispatcher ("no kernel for CPU"), never falls back.

The descriptor travels as a list of ints + a list of floats (``b200a_frontend_desc`` is plain old data).
"""
from __future__ import annotations

from typing import List, Optional

import torch

from . import _lib

_DESC_INTS = ("n_fft", "win_length", "hop", "pad", "center", "pad_mode", "onesided", "frame_length_norm", "window_norm",
              "n_mels", "n_mfcc", "log_mels")
_DESC_FLOATS = ("power", "db_multiplier", "db_amin", "db_offset")

_LIB = torch.library.Library("b200audio", "DEF")
_LIB.define(
    "frontend_run(Tensor wave, Tensor workspace, int[] desc_i, float[] desc_f, int stage, int frames, int width, "
    "int row_stride, Tensor(a!)? group_max, int rows_per_group) -> Tensor"
)
_LIB.define(
    "mfcc_finish(Tensor feat, Tensor workspace, int[] desc_i, float[] desc_f, Tensor? group_max, int rows_per_group, "
    "float top_db) -> Tensor"
)
_LIB.define(
    "resample_run(Tensor wave, Tensor workspace, Tensor kernel, int orig_r, int new_r, int width, int row_stride, "
    "int out_len, int pitch) -> Tensor"
)


def pack_desc(d: "_lib.FrontendDesc"):
    return [int(getattr(d, k)) for k in _DESC_INTS], [float(getattr(d, k)) for k in _DESC_FLOATS]


def _unpack_desc(desc_i: List[int], desc_f: List[float]) -> "_lib.FrontendDesc":
    d = _lib.FrontendDesc()
    for k, v in zip(_DESC_INTS, desc_i):
        setattr(d, k, int(v))
    for k, v in zip(_DESC_FLOATS, desc_f):
        setattr(d, k, float(v))
    return d


def _stream(dev: torch.device) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def _out_shape(wave, stage, frames, width):
    rows = wave.shape[0]
    return (rows, frames, width, 2) if stage == _lib.STAGE_COMPLEX else (rows, frames, width)


# ---- frontend_run ------------------------------------------------------------------------------------------------
def _frontend_run_cuda(wave, workspace, desc_i, desc_f, stage, frames, width, row_stride, group_max, rows_per_group):
    d = _unpack_desc(desc_i, desc_f)
    dev = wave.device
    with torch.cuda.device(dev):
        out = torch.empty(_out_shape(wave, stage, frames, width), dtype=torch.float32, device=dev)
        rc = _lib.lib().b200a_frontend_run(
            d, workspace.data_ptr(), stage, wave.data_ptr(), wave.shape[0], wave.shape[1], row_stride, out.data_ptr(),
            None if group_max is None else group_max.data_ptr(), rows_per_group, _stream(dev))
    if rc == _lib.ESHORT:
        raise RuntimeError(
            f"audio_b200: padding size n_fft//2={d.n_fft // 2} should be less than the input length "
            f"{wave.shape[1] + 2 * d.pad} for pad_mode reflect/circular (torch.stft raises the same way)")
    _lib.check(rc, "frontend_run")
    return out


def _frontend_run_meta(wave, workspace, desc_i, desc_f, stage, frames, width, row_stride, group_max, rows_per_group):
    return wave.new_empty(_out_shape(wave, stage, frames, width), dtype=torch.float32)


# ---- mfcc_finish -------------------------------------------------------------------------------------------------
def _mfcc_finish_cuda(feat, workspace, desc_i, desc_f, group_max, rows_per_group, top_db):
    d = _unpack_desc(desc_i, desc_f)
    rows, frames, _ = feat.shape
    dev = feat.device
    with torch.cuda.device(dev):
        out = torch.empty((rows, frames, d.n_mfcc), dtype=torch.float32, device=dev)
        rc = _lib.lib().b200a_mfcc_finish(
            d, workspace.data_ptr(), feat.data_ptr(), rows, frames, None if group_max is None else group_max.data_ptr(),
            rows_per_group, float(top_db), out.data_ptr(), _stream(dev))
    _lib.check(rc, "mfcc_finish")
    return out


def _mfcc_finish_meta(feat, workspace, desc_i, desc_f, group_max, rows_per_group, top_db):
    return feat.new_empty((feat.shape[0], feat.shape[1], int(desc_i[_DESC_INTS.index("n_mfcc")])))


# ---- resample_run ------------------------------------------------------------------------------------------------
def _resample_run_cuda(wave, workspace, kernel, orig_r, new_r, width, row_stride, out_len, pitch):
    rows, length = wave.shape
    dev = wave.device
    with torch.cuda.device(dev):
        buf = torch.empty((rows, pitch), dtype=torch.float32, device=dev)
        rc = _lib.lib().b200a_resample_run(
            workspace.data_ptr(), kernel.data_ptr(), orig_r, new_r, width, wave.data_ptr(), rows, length, row_stride,
            buf.data_ptr(), pitch, out_len, _stream(dev))
    _lib.check(rc, "resample_run")
    return buf


def _resample_run_meta(wave, workspace, kernel, orig_r, new_r, width, row_stride, out_len, pitch):
    return wave.new_empty((wave.shape[0], pitch))


for _name, _cuda, _meta in (("frontend_run", _frontend_run_cuda, _frontend_run_meta),
                            ("mfcc_finish", _mfcc_finish_cuda, _mfcc_finish_meta),
                            ("resample_run", _resample_run_cuda, _resample_run_meta)):
    _LIB.impl(_name, _cuda, "CUDA")
    _LIB.impl(_name, _meta, "Meta")

frontend_run = torch.ops.b200audio.frontend_run
mfcc_finish = torch.ops.b200audio.mfcc_finish
resample_run = torch.ops.b200audio.resample_run
