"""Device-side plans: a descriptor + the caller-owned workspace libb200audio fills once.

PyTorch is used here only for what the C ABI deliberately leaves to the caller: device
memory (``torch.empty``), the current stream, and the device guard.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch

from . import _lib, _ops
from ._bookkeeping import resample_len


def _require_cuda_f32(t: torch.Tensor, what: str) -> None:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{what} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(
            f"audio_b200: {what} is on '{t.device}'. This package runs only hand-written sm_100a CUDA "
            "kernels; there is no CPU or ATen fallback -- move the tensor (and the module) to a CUDA device."
        )
    if t.dtype != torch.float32:
        raise TypeError(f"audio_b200: {what} must be float32 (got {t.dtype}); other dtypes are not implemented")


def _no_autograd(t: torch.Tensor) -> None:
    if t.requires_grad and torch.is_grad_enabled():
        raise RuntimeError(
            "audio_b200 kernels are forward-only: the input requires grad. Call under torch.no_grad() / "
            "torch.inference_mode(), or detach() the input."
        )


def _version_of(t: torch.Tensor) -> int:
    """In-place edit counter of a tensor; inference tensors (created under torch.inference_mode) keep none."""
    return -1 if t.is_inference() else t._version


def _stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def pack_rows(waveform: torch.Tensor) -> Tuple[torch.Tensor, int]:
    """(..., time) -> (rows, time) view with unit inner stride; returns it and the row stride."""
    length = waveform.shape[-1]
    flat = waveform.reshape(-1, length)
    if flat.shape[0] > 0 and length > 0 and (flat.stride(1) != 1 or (flat.shape[0] > 1 and flat.stride(0) < length)):
        flat = flat.contiguous()
    stride = flat.stride(0) if flat.shape[0] > 1 else max(length, 1)
    return flat, stride


class FrontendPlan:
    """Workspace for one (descriptor, window, fb, dct) combination on one device."""

    def __init__(self, desc: "_lib.FrontendDesc"):
        self.desc = desc
        self._ws: Optional[torch.Tensor] = None
        self._stamp = None
        # the constant tensors the workspace was built from: holding them keeps their storage from being
        # recycled, so an equal (data_ptr, _version) stamp can only mean "the same, unmodified tensor"
        self._held = None
        self._desc_lists = None  # the descriptor as (ints, floats) for the torch.library ops

    @staticmethod
    def make_desc(
        n_fft: int,
        win_length: int,
        hop: int,
        pad: int,
        center: bool,
        pad_mode: str,
        onesided: bool,
        frame_length_norm: bool,
        window_norm: bool,
        power: Optional[float],
        n_mels: int = 0,
        n_mfcc: int = 0,
        log_mels: bool = False,
    ) -> "_lib.FrontendDesc":
        if pad_mode not in _lib.PAD_MODE:
            raise ValueError(f"Unsupported pad_mode: {pad_mode!r} (expected one of {sorted(_lib.PAD_MODE)})")
        d = _lib.FrontendDesc()
        d.n_fft, d.win_length, d.hop, d.pad = int(n_fft), int(win_length), int(hop), int(pad)
        d.center, d.pad_mode, d.onesided = int(bool(center)), _lib.PAD_MODE[pad_mode], int(bool(onesided))
        d.frame_length_norm, d.window_norm = int(bool(frame_length_norm)), int(bool(window_norm))
        d.power = float("nan") if power is None else float(power)
        d.n_mels, d.n_mfcc, d.log_mels = int(n_mels), int(n_mfcc), int(bool(log_mels))
        d.db_multiplier, d.db_amin = 10.0, 1e-10
        d.db_offset = 10.0 * math.log10(max(1e-10, 1.0))
        return d

    def _stamp_of(self, *tensors):
        return tuple(None if t is None else (t.data_ptr(), _version_of(t), str(t.device)) for t in tensors)

    def workspace(self, window: torch.Tensor, fb: Optional[torch.Tensor], dct: Optional[torch.Tensor]) -> torch.Tensor:
        """Return a prepared workspace, rebuilding it if any constant buffer changed."""
        stamp = self._stamp_of(window, fb, dct)
        if self._ws is not None and stamp == self._stamp:
            return self._ws
        lib = _lib.lib()
        _require_cuda_f32(window, "window")
        dev = window.device
        for name, t in (("fb", fb), ("dct_mat", dct)):
            if t is not None:
                _require_cuda_f32(t, name)
                if t.device != dev:
                    raise RuntimeError(f"audio_b200: {name} is on {t.device} but window is on {dev}")
        if window.numel() != self.desc.win_length:
            raise RuntimeError(
                f"expected a 1D window tensor of size equal to win_length={self.desc.win_length}, "
                f"but got: window size={window.numel()}"
            )
        nbytes = lib.b200a_frontend_workspace_bytes(self.desc)
        if nbytes == 0:
            raise _lib.B200AudioError(_lib.EUNSUPPORTED, "frontend_workspace_bytes (descriptor rejected)")
        window_c = window.contiguous()
        fb_c = None if fb is None else fb.contiguous()
        dct_c = None if dct is None else dct.contiguous()
        with torch.cuda.device(dev):
            ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            rc = lib.b200a_frontend_prepare(
                self.desc,
                window_c.data_ptr(),
                None if fb_c is None else fb_c.data_ptr(),
                None if dct_c is None else dct_c.data_ptr(),
                ws.data_ptr(),
                nbytes,
                _stream_ptr(dev),
            )
        _lib.check(rc, "frontend_prepare")
        self._ws, self._stamp, self._held = ws, stamp, (window, fb, dct)
        return ws

    def frames(self, length: int) -> int:
        return _lib.lib().b200a_num_frames(length, self.desc.n_fft, self.desc.hop, self.desc.center, self.desc.pad)

    def run(
        self,
        ws: torch.Tensor,
        stage: int,
        waveform: torch.Tensor,
        group_max: Optional[torch.Tensor] = None,
        rows_per_group: int = 1,
    ) -> torch.Tensor:
        """Launch the fused kernel; returns the FRAME-MAJOR result (rows, T, width[, 2])."""
        _require_cuda_f32(waveform, "waveform")
        _no_autograd(waveform)
        if waveform.device != ws.device:
            raise RuntimeError(f"audio_b200: waveform is on {waveform.device} but the module buffers are on {ws.device}")
        lib = _lib.lib()
        d = self.desc
        flat, stride = pack_rows(waveform)
        rows, length = flat.shape
        frames = self.frames(length)
        if frames < 1:
            raise RuntimeError(
                f"audio_b200: waveform of {length} samples is too short for n_fft={d.n_fft} "
                f"(center={bool(d.center)}, pad={d.pad})"
            )
        n_bins = lib.b200a_num_bins(d.n_fft, d.onesided)
        width = d.n_mels if stage >= _lib.STAGE_MEL else n_bins
        # through the dispatcher (b200audio::frontend_run, audio_b200/_ops.py): allocates `out`, launches on the
        # current stream of the waveform's device, raises on a negative status
        desc_i, desc_f = self._packed_desc()
        return _ops.frontend_run(flat, ws, desc_i, desc_f, stage, frames, width, stride, group_max, rows_per_group)

    def mfcc_finish(self, ws, feat, group_max, rows_per_group: int, top_db: Optional[float]) -> torch.Tensor:
        desc_i, desc_f = self._packed_desc()
        return _ops.mfcc_finish(feat, ws, desc_i, desc_f, group_max, rows_per_group, -1.0 if top_db is None else float(top_db))

    def _packed_desc(self):
        if self._desc_lists is None:
            self._desc_lists = _ops.pack_desc(self.desc)
        return self._desc_lists


def new_group_max(groups: int, device: torch.device) -> torch.Tensor:
    """[groups] running maxima initialised to -inf by the library's fill kernel."""
    with torch.cuda.device(device):
        g = torch.empty(groups, dtype=torch.float32, device=device)
        _lib.check(_lib.lib().b200a_fill_f32(g.data_ptr(), groups, float("-inf"), _stream_ptr(device)), "fill_f32")
    return g


class ResamplePlan:
    """Per-phase tap supports of a cached sinc kernel, kept next to the kernel buffer."""

    def __init__(self, orig_r: int, new_r: int, width: int):
        self.orig_r, self.new_r, self.width = int(orig_r), int(new_r), int(width)
        self.taps = 2 * self.width + self.orig_r
        self._ws: Optional[torch.Tensor] = None
        self._stamp = None
        self._kernel: Optional[torch.Tensor] = None
        self._held = None  # the kernel tensor behind the stamp (see FrontendPlan._held)

    def workspace(self, kernel: torch.Tensor):
        stamp = (kernel.data_ptr(), _version_of(kernel), str(kernel.device))
        if self._ws is not None and stamp == self._stamp:
            return self._ws, self._kernel
        _require_cuda_f32(kernel, "kernel")
        if kernel.numel() != self.new_r * self.taps:
            raise RuntimeError(
                f"audio_b200: resample kernel has {kernel.numel()} elements, expected {self.new_r}x{self.taps}"
            )
        lib = _lib.lib()
        dev = kernel.device
        k = kernel.reshape(self.new_r, self.taps).contiguous()
        nbytes = lib.b200a_resample_workspace_bytes(self.new_r, self.taps)
        with torch.cuda.device(dev):
            ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            rc = lib.b200a_resample_prepare(k.data_ptr(), self.orig_r, self.new_r, self.width, ws.data_ptr(), nbytes, _stream_ptr(dev))
        _lib.check(rc, "resample_prepare")
        self._ws, self._stamp, self._kernel, self._held = ws, stamp, k, kernel
        return ws, k

    def run(self, kernel: torch.Tensor, waveform: torch.Tensor) -> torch.Tensor:
        _require_cuda_f32(waveform, "waveform")
        _no_autograd(waveform)
        ws, k = self.workspace(kernel)
        if waveform.device != ws.device:
            raise RuntimeError(f"audio_b200: waveform is on {waveform.device} but the kernel buffer is on {ws.device}")
        lib = _lib.lib()
        flat, stride = pack_rows(waveform)
        rows, length = flat.shape
        out_len = resample_len(length, self.orig_r, self.new_r)
        # the reference returns a view into (rows, frames*new') memory: keep that row pitch
        pitch = (length // self.orig_r + 1) * self.new_r
        buf = _ops.resample_run(flat, ws, k, self.orig_r, self.new_r, self.width, stride, out_len, pitch)
        out = buf[:, :out_len]
        return out.view(waveform.shape[:-1] + (out_len,)) if rows > 0 else out.reshape(waveform.shape[:-1] + (out_len,))
