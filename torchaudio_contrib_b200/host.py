"""Host-buffer entry: CPU tensors in, CPU tensors out, compute on the B200 (tac_pipeline_*).

This is what a caller holding host memory binds -- the reference's CPU-runnable configuration
(`Spectrogram(fft_length=512, hop_length=128)` on a CPU tensor) goes through here.  It is not a
CPU fallback: the library copies the batch to the device in slices, runs the same kernels and
copies the result back, overlapping both copy directions with compute (one stream each for H2D, kernels, D2H).
"""
import ctypes

import torch

from . import _cabi
from . import functional as F


class HostPipeline(object):
    """`Spectrogram` (filterbank=None) or `Melspectrogram[+AmplitudeToDb]` on host buffers.

    Arguments follow `STFT` / `Melspectrogram` / `AmplitudeToDb`.  `__call__` takes a float32
    CPU tensor `(*, channel, time)` (pinned memory makes the copies asynchronous) and returns a
    CPU tensor `(*, channel, num_bands | num_freqs, frames)`.
    """

    def __init__(self, fft_length, hop_length=None, win_length=None, window=None, center=True, pad_mode='reflect',
                 normalized=False, power=1.0, filterbank=None, to_db=False, ref=1.0, amin=1e-7, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("HostPipeline needs a CUDA device (there is no CPU implementation)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.fft_length = int(fft_length)
        self.hop = self.fft_length // 4 if hop_length is None else int(hop_length)
        self.center = bool(center)
        win = F._frame_window(window, win_length, self.fft_length, "cpu")
        cfg = _cabi.PipelineConfig()
        cfg.n_fft, cfg.hop, cfg.center = self.fft_length, self.hop, int(self.center)
        cfg.pad_mode, cfg.normalized, cfg.power = _cabi.PAD_MODES[pad_mode], int(bool(normalized)), float(power)
        cfg.to_db, cfg.ref, cfg.amin = int(bool(to_db)), float(ref), float(amin)
        fb = None
        if filterbank is not None:
            fb = filterbank.detach().to(device="cpu", dtype=torch.float32).contiguous()
            cfg.n_bins, cfg.n_bands = int(fb.size(0)), int(fb.size(1))
        self.rows = cfg.n_bands if fb is not None else self.fft_length // 2 + 1
        self._handle = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _cabi.check(_cabi.lib().tac_pipeline_create(ctypes.byref(cfg), _cabi.ptr(win),
                                                        _cabi.ptr(fb) if fb is not None else None,
                                                        ctypes.byref(self._handle)))

    def frames(self, n_samples):
        return int(_cabi.lib().tac_stft_num_frames(n_samples, self.fft_length, self.hop, int(self.center)))

    def __call__(self, waveforms, out=None):
        if waveforms.is_cuda or waveforms.dtype != torch.float32:
            raise RuntimeError("HostPipeline expects a float32 CPU tensor")
        x = waveforms.contiguous()
        lead, n_samples = x.shape[:-1], x.size(-1)
        n_seq = x.numel() // max(n_samples, 1)
        shape = tuple(lead) + (self.rows, self.frames(n_samples))
        if out is None:
            out = torch.empty(shape, dtype=torch.float32)
        elif tuple(out.shape) != shape or not out.is_contiguous() or out.dtype != torch.float32:
            raise RuntimeError("HostPipeline: `out` must be a contiguous float32 tensor of shape %s" % (shape,))
        with torch.cuda.device(self.device):               # the library selects its device; restore the caller's afterwards
            _cabi.check(_cabi.lib().tac_pipeline_run_host(self._handle, _cabi.ptr(x), n_seq, n_samples, _cabi.ptr(out)))
        return out

    def close(self):
        if getattr(self, "_handle", None):
            _cabi.lib().tac_pipeline_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
