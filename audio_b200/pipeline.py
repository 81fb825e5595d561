"""Host-buffer pipeline: run a front-end module over waveforms that live in (pinned) HOST memory.

A B200 turns 256 x 10 s of audio into mel features in a fraction of a millisecond, but the same
batch takes ~3 ms to cross PCIe.  ``HostPipeline`` splits the batch into row chunks and keeps three
CUDA streams busy -- host->device copies, the fused kernel, device->host copies -- so the end-to-end
time approaches the slower of the two PCIe directions instead of their sum plus the compute.

    pipe = HostPipeline(T.MelSpectrogram(16000, n_fft=1024, hop_length=256, n_mels=80).cuda())
    feats = pipe(wave_host)              # (B, n_mels, T) view of a pinned frame-major (B, T, n_mels) buffer
    torch.cuda.synchronize()             # or pipe.synchronize()

Rows are independent on this path, so chunking never changes results -- with ONE
exception the caller must respect: ``MFCC`` on a 2-D batch shares one top_db maximum over the whole
batch (reference functional.py:395-399); chunking such a call would change it, so it is refused.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import transforms as T


class HostPipeline:
    def __init__(self, module: torch.nn.Module, chunk_rows: int = 32, device: Optional[torch.device] = None):
        if isinstance(module, T.MFCC) and not module.log_mels:
            raise ValueError(
                "HostPipeline cannot chunk MFCC with the batch-global top_db clamp; pass (B, 1, time) inputs "
                "through the module directly or use log_mels=True"
            )
        params = list(module.buffers())
        self.device = device if device is not None else (params[0].device if params else torch.device("cuda"))
        if self.device.type != "cuda":
            raise RuntimeError("HostPipeline needs the module on a CUDA device (no CPU fallback)")
        self.module = module
        self.chunk_rows = int(chunk_rows)
        self._s_in = torch.cuda.Stream(self.device)
        self._s_run = torch.cuda.Stream(self.device)
        self._s_out = torch.cuda.Stream(self.device)
        self._dev_in = None
        self._host_out = None

    def synchronize(self) -> None:
        self._s_out.synchronize()

    @torch.inference_mode()
    def __call__(self, wave_host: torch.Tensor, out_host: Optional[torch.Tensor] = None) -> torch.Tensor:
        if wave_host.is_cuda or wave_host.dim() != 2 or wave_host.dtype != torch.float32:
            raise TypeError("HostPipeline expects a 2-D float32 CPU tensor (pin it for asynchronous copies)")
        rows, length = wave_host.shape
        step = max(1, min(self.chunk_rows, rows))
        n_chunks = (rows + step - 1) // step
        if self._dev_in is None or self._dev_in.shape[1:] != (step, length):
            with torch.cuda.device(self.device):
                self._dev_in = torch.empty((3, step, length), dtype=torch.float32, device=self.device)
        ev_in = [torch.cuda.Event() for _ in range(n_chunks)]
        ev_run = [torch.cuda.Event() for _ in range(n_chunks)]
        caller = torch.cuda.current_stream(self.device)
        for s in (self._s_in, self._s_run, self._s_out):
            s.wait_stream(caller)
        result = out_host
        for i in range(n_chunks):
            lo, hi = i * step, min(rows, (i + 1) * step)
            buf = self._dev_in[i % 3, : hi - lo]
            with torch.cuda.stream(self._s_in):
                if i >= 3:
                    self._s_in.wait_event(ev_run[i - 3])  # the kernel that read this buffer has finished
                buf.copy_(wave_host[lo:hi], non_blocking=True)
                ev_in[i].record(self._s_in)
            with torch.cuda.stream(self._s_run):
                self._s_run.wait_event(ev_in[i])
                y = self.module(buf)  # logical (rows, W, T) view of frame-major (rows, T, W) memory
                ev_run[i].record(self._s_run)
            y_fm = y.transpose(-1, -2)
            if result is None:
                result = torch.empty((rows,) + tuple(y_fm.shape[1:]), dtype=y.dtype).pin_memory()
            with torch.cuda.stream(self._s_out):
                self._s_out.wait_event(ev_run[i])
                y_fm.record_stream(self._s_out)
                result[lo:hi].copy_(y_fm, non_blocking=True)
        caller.wait_stream(self._s_out)
        self._host_out = result
        return result.transpose(-1, -2)
