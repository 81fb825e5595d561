"""ctypes binding of libb200audio.so (the C ABI in include/b200audio.h).

This is the stub a pytorch/audio maintainer would add next to
``src/torchaudio/_extension/utils.py:_load_lib``: load the shared object, declare the
argument types, turn negative status codes into exceptions.  There is NO fallback: if the
library is missing the import of any op raises, and ops refuse non-CUDA tensors.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int32, c_int64, c_size_t, c_void_p

_PKG = os.path.dirname(os.path.abspath(__file__))
# B200A_LIB: load another build of the same ABI (A/B timing of kernel changes); the default is the in-tree build
LIB_PATH = os.environ.get("B200A_LIB") or os.path.join(_PKG, "lib", "libb200audio.so")

OK, EINVAL, EUNSUPPORTED, ESHORT, EWORKSPACE, ECUDA = 0, -1, -2, -3, -4, -5
PAD_MODE = {"reflect": 0, "constant": 1, "replicate": 2, "circular": 3}
STAGE_COMPLEX, STAGE_POWER, STAGE_MEL, STAGE_FEAT = 0, 1, 2, 3


class FrontendDesc(ctypes.Structure):
    """Mirror of ``b200a_frontend_desc``."""

    _fields_ = [
        ("n_fft", c_int32),
        ("win_length", c_int32),
        ("hop", c_int32),
        ("pad", c_int32),
        ("center", c_int32),
        ("pad_mode", c_int32),
        ("onesided", c_int32),
        ("frame_length_norm", c_int32),
        ("window_norm", c_int32),
        ("power", c_float),
        ("n_mels", c_int32),
        ("n_mfcc", c_int32),
        ("log_mels", c_int32),
        ("db_multiplier", c_float),
        ("db_amin", c_float),
        ("db_offset", c_float),
    ]

    def key(self):
        return tuple(getattr(self, f) for f, _ in self._fields_)


class KaldiDesc(ctypes.Structure):
    """Mirror of ``b200a_kaldi_desc``."""

    _fields_ = [
        ("window_size", c_int32),
        ("window_shift", c_int32),
        ("padded_size", c_int32),
        ("snip_edges", c_int32),
        ("remove_dc_offset", c_int32),
        ("preemphasis", c_float),
        ("energy_mode", c_int32),
        ("energy_floor", c_float),
        ("energy_col", c_int32),
        ("out_width", c_int32),
        ("out_col0", c_int32),
        ("use_log", c_int32),
    ]


_SIGNATURES = {
    "b200a_version": (ctypes.c_int, []),
    "b200a_strerror": (c_char_p, [ctypes.c_int]),
    "b200a_num_frames": (c_int64, [c_int64, c_int32, c_int32, c_int32, c_int32]),
    "b200a_pad_index": (c_int64, [c_int64, c_int64, c_int32]),
    "b200a_num_bins": (c_int32, [c_int32, c_int32]),
    "b200a_resample_width": (c_int32, [c_int32, c_int32, c_int32, c_double]),
    "b200a_resample_len": (c_int64, [c_int64, c_int32, c_int32]),
    "b200a_resample_support": (ctypes.c_int, [c_int32, c_int32, c_int32, c_double, c_int32, POINTER(c_int32), POINTER(c_int32)]),
    "b200a_resample_plan_info": (ctypes.c_int, [c_int32, c_int32, c_int32, POINTER(c_int32)]),
    "b200a_resample_tc_band": (ctypes.c_int, [c_int32, c_int32, c_int32, c_int32, POINTER(c_int32), POINTER(c_int32)]),
    "b200a_frontend_workspace_bytes": (c_size_t, [POINTER(FrontendDesc)]),
    "b200a_frontend_prepare": (
        ctypes.c_int,
        [POINTER(FrontendDesc), c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p],
    ),
    "b200a_frontend_run": (
        ctypes.c_int,
        [POINTER(FrontendDesc), c_void_p, c_int32, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_int64, c_void_p],
    ),
    "b200a_mfcc_finish": (
        ctypes.c_int,
        [POINTER(FrontendDesc), c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_int64, c_float, c_void_p, c_void_p],
    ),
    "b200a_apply_fbank": (
        ctypes.c_int,
        [c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_void_p, c_int32, c_void_p, c_void_p],
    ),
    "b200a_amplitude_to_db": (
        ctypes.c_int,
        [c_void_p, c_int64, c_int64, c_float, c_float, c_float, c_float, c_void_p, c_void_p, c_void_p],
    ),
    "b200a_istft_run": (
        ctypes.c_int,
        [POINTER(FrontendDesc), c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_int64,
         c_int64, c_int64, c_void_p],
    ),
    "b200a_griffinlim_update": (
        ctypes.c_int,
        [c_void_p, c_int64, c_int64, c_int64, c_float, c_void_p, c_void_p, c_float, c_int32, c_void_p, c_int64, c_int64, c_int64,
         c_void_p],
    ),
    "b200a_phase_vocoder": (
        ctypes.c_int,
        [c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_double, c_void_p, c_void_p, c_int64, c_void_p],
    ),
    "b200a_kaldi_num_frames": (c_int64, [c_int64, c_int32, c_int32, c_int32]),
    "b200a_kaldi_run": (
        ctypes.c_int,
        [POINTER(KaldiDesc), POINTER(FrontendDesc), c_void_p, c_int32, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p],
    ),
    "b200a_subtract_column_mean": (ctypes.c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p]),
    "b200a_fill_f32": (ctypes.c_int, [c_void_p, c_int64, c_float, c_void_p]),
    "b200a_ratio_f32": (ctypes.c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    "b200a_resample_workspace_bytes": (c_size_t, [c_int32, c_int32]),
    "b200a_resample_prepare": (ctypes.c_int, [c_void_p, c_int32, c_int32, c_int32, c_void_p, c_size_t, c_void_p]),
    "b200a_resample_run": (
        ctypes.c_int,
        [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int64, c_int64, c_void_p],
    ),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


class B200AudioError(RuntimeError):
    def __init__(self, status: int, where: str):
        self.status = status
        msg = lib().b200a_strerror(status).decode()
        super().__init__(f"libb200audio: {where}: {msg} (status {status})")


def lib() -> ctypes.CDLL:
    """Load (once) and return the shared library; raises if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing. Build it with `python -m audio_b200._build` "
                "(nvcc, sm_100a). audio_b200 has no CPU or ATen fallback."
            )
        handle = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in _SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError here == header/library mismatch
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib


def check(status: int, where: str) -> None:
    if status != OK:
        raise B200AudioError(status, where)
