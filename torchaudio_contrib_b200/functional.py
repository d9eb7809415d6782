"""Functional surface of torchaudio_contrib (reference: torchaudio_contrib/functional.py) on
hand-written sm_100a kernels.

Same names, argument meaning and shapes as the reference; every function documents the reference
lines it stands in for.  Tensors must live on a CUDA device and be float32 (the tuned kernels) or float64 (plain double
kernels, forward only; mu-law codes int64): each call enqueues kernels of libtac_b200.so on the current stream.  There is
no CPU path.

Autograd: the signal path (stft, complex_norm, apply_filterbank, amplitude_to_db, spectrogram, melspectrogram,
db_to_amplitude, magphase / angle, float-input mu_law_decoding) is differentiated by hand-written adjoint kernels,
and the `filterbank` and `window` arguments get their gradients too (a learnable filterbank / window).  What is not differentiated raises
instead of silently detaching: `mu_law_encoding`
inputs that require grad and float64 inputs (forward only, except `phase_vocoder`).  `phase_vocoder` is differentiated
w.r.t. the spectrogram by its own gather kernel; every `fft_length` in [2, 8192] differentiates.
"""
import collections
import ctypes
import math
import os
import threading

import torch

from . import _cabi, _f64, _mulaw_tables

__all__ = [
    "stft", "complex_norm", "create_mel_filter", "apply_filterbank", "amplitude_to_db",
    "mu_law_encoding", "mu_law_decoding", "spectrogram", "melspectrogram", "FilterbankPlan",
    "PreparedMelspectrogram", "angle", "magphase", "phase_vocoder", "db_to_amplitude",
]


# ------------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------------
def _wants_grad(t):
    return torch.is_grad_enabled() and isinstance(t, torch.Tensor) and t.requires_grad


def _no_param_grad(t, name):
    if isinstance(t, torch.Tensor) and torch.is_grad_enabled() and t.requires_grad:
        raise RuntimeError("%s requires grad, but it is a constant on this path (the adjoint kernels differentiate w.r.t. the "
                           "signal, the filterbank and the float32 window)" % name)


def _forward_only(t, name):
    if torch.is_grad_enabled() and t.requires_grad:
        raise RuntimeError("%s: the B200 kernels are forward-only; call under torch.no_grad() or detach "
                           "the input (the reference is differentiable through torch, this path is not)" % name)


def _as_f32_cuda(t, name):
    _cabi.require_cuda(t, name)
    if t.dtype != torch.float32:
        raise NotImplementedError("%s has dtype %s: float32 (tuned kernels) and float64 (csrc/f64_path.cu) are implemented"
                                  % (name, t.dtype))
    return t.contiguous()


_default_windows = {}


def _frame_window(window, win_length, fft_length, device):
    """The n_fft-long window torch.stft effectively multiplies by: Hann(win_length) when none is
    given (functional.py:93-97), zero-padded on both sides to sit in the middle of the frame."""
    if win_length is None:
        win_length = fft_length
    if window is None:
        key = (win_length, fft_length, str(device))
        cached = _default_windows.get(key)
        if cached is not None:
            return cached
        # evaluated on the host like the reference's (functional.py:93-97, layers.py:76-82): the CUDA
        # hann_window differs from it in the last bit, which would make functional and module outputs differ
        window = torch.hann_window(win_length)
    else:
        key = None
    if window.dim() != 1 or window.size(0) != win_length:
        raise RuntimeError("stft: expected a 1-D window of size win_length=%d, got %s"
                           % (win_length, tuple(window.shape)))
    if not (0 < win_length <= fft_length):
        raise RuntimeError("stft: expected 0 < win_length <= n_fft, got win_length=%d n_fft=%d"
                           % (win_length, fft_length))
    window = window.to(device=device, dtype=torch.float32)
    if win_length < fft_length:
        left = (fft_length - win_length) // 2
        window = torch.nn.functional.pad(window, (left, fft_length - win_length - left))
    window = window.contiguous()
    if key is not None and 0 < win_length <= fft_length:
        _default_windows[key] = window
    return window


def _stft_geometry(waveforms, fft_length, hop_length, center):
    hop = fft_length // 4 if hop_length is None else int(hop_length)
    lead = waveforms.shape[:-1]
    n_samples = waveforms.size(-1)
    flat = waveforms.reshape(-1, n_samples)
    frames = int(_cabi.lib().tac_stft_num_frames(n_samples, fft_length, hop, int(bool(center))))
    return hop, lead, flat, frames


def _stft_args(flat, window, fft_length, hop, center, pad_mode, normalized):
    if pad_mode not in _cabi.PAD_MODES:
        raise NotImplementedError("stft: pad_mode=%r (supported: %s)" % (pad_mode, sorted(_cabi.PAD_MODES)))
    return [_cabi.ptr(flat), flat.size(0), flat.size(1), flat.stride(0) if flat.size(0) > 1 else flat.size(1),
            _cabi.ptr(window), int(fft_length), hop, int(bool(center)), _cabi.PAD_MODES[pad_mode],
            int(bool(normalized))]


# ------------------------------------------------------------------------------------------------
# a1: stft
# ------------------------------------------------------------------------------------------------
def stft(waveforms, fft_length, hop_length=None, win_length=None, window=None,
         center=True, pad_mode='reflect', normalized=False, onesided=True):
    """Short-time Fourier transform, `(*, channel, time) -> (*, channel, num_freqs, frames, 2)`.

    Reference: functional.py:48-113 (a reshape around `torch.stft`, :99-107).  Here: one fused
    framing + padding + window + real-FFT kernel (csrc/stft.cu) for powers of two in [32, 8192], a
    direct-DFT kernel for any other `fft_length` in [2, 8192]; float64 waveforms take the double
    kernels of csrc/f64_path.cu (forward only).  Unlike the reference, a missing `window` is created
    on the input's device.  The result is a contiguous tensor of the reference's logical shape.
    """
    if isinstance(waveforms, torch.Tensor) and waveforms.dtype == torch.float64:
        _no_param_grad(window, "stft (float64): window")
        return _f64.stft(waveforms, fft_length, hop_length, win_length, window, center, pad_mode, normalized, onesided)
    if _wants_grad(waveforms) or _wants_grad(window):
        return _StftFn.apply(waveforms, window if _wants_grad(window) else None,
                             (fft_length, hop_length, win_length, window, center, pad_mode, normalized, onesided))
    x = _as_f32_cuda(waveforms, "waveforms")
    hop, lead, flat, frames = _stft_geometry(x, fft_length, hop_length, center)
    win = _frame_window(window, win_length, fft_length, x.device)
    bins = fft_length // 2 + 1 if onesided else fft_length
    out = torch.empty((flat.size(0), bins, max(frames, 0), 2), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _cabi.check(_cabi.lib().tac_stft_f32(
            *_stft_args(flat, win, fft_length, hop, center, pad_mode, normalized),
            int(bool(onesided)), _cabi.ptr(out), _cabi.stream_ptr(x.device)))
    return out.reshape(lead + out.shape[1:])


def spectrogram(waveforms, fft_length, hop_length=None, win_length=None, window=None, center=True,
                pad_mode='reflect', normalized=False, onesided=True, power=1.):
    """`Spectrogram(...)(x)` in one kernel: stft then `|.|^power` (layers.py:294-304), the complex
    spectrum never reaches HBM.  Returns `(*, channel, num_freqs, frames)`."""
    if isinstance(waveforms, torch.Tensor) and waveforms.dtype == torch.float64:
        _no_param_grad(window, "spectrogram (float64): window")
        return _f64.complex_norm(_f64.stft(waveforms, fft_length, hop_length, win_length, window, center, pad_mode, normalized,
                                           onesided), float(power))
    if _wants_grad(window):        # a learnable window: the stages, each with its own adjoint kernels (the fused adjoint treats it as a constant)
        return complex_norm(stft(waveforms, fft_length, hop_length, win_length, window, center, pad_mode, normalized, onesided), power)
    if _wants_grad(waveforms):
        return _SpectrogramFn.apply(waveforms, (fft_length, hop_length, win_length, window, center, pad_mode, normalized,
                                                onesided, power))
    x = _as_f32_cuda(waveforms, "waveforms")
    hop, lead, flat, frames = _stft_geometry(x, fft_length, hop_length, center)
    win = _frame_window(window, win_length, fft_length, x.device)
    bins = fft_length // 2 + 1 if onesided else fft_length
    out = torch.empty((flat.size(0), bins, max(frames, 0)), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _cabi.check(_cabi.lib().tac_spectrogram_f32(
            *_stft_args(flat, win, fft_length, hop, center, pad_mode, normalized),
            int(bool(onesided)), float(power), _cabi.ptr(out), _cabi.stream_ptr(x.device)))
    return out.reshape(lead + out.shape[1:])


# ------------------------------------------------------------------------------------------------
# a2: complex_norm
# ------------------------------------------------------------------------------------------------
def complex_norm(complex_tensor, power=1.0):
    """`(*, 2) -> (*)`: sqrt(re^2 + im^2), then `.pow(power)` (functional.py:116-128)."""
    if isinstance(complex_tensor, torch.Tensor) and complex_tensor.dtype == torch.float64:
        return _f64.complex_norm(complex_tensor, float(power))
    if _wants_grad(complex_tensor):
        return _ComplexNormFn.apply(complex_tensor, float(power))
    z = _as_f32_cuda(complex_tensor, "complex_tensor")
    if z.dim() < 1 or z.size(-1) != 2:
        raise RuntimeError("complex_norm: expected a (*, 2) tensor, got %s" % (tuple(z.shape),))
    out = torch.empty(z.shape[:-1], dtype=torch.float32, device=z.device)
    with torch.cuda.device(z.device):
        _cabi.check(_cabi.lib().tac_complex_norm_f32(_cabi.ptr(z), out.numel(), float(power), _cabi.ptr(out),
                                                     _cabi.stream_ptr(z.device)))
    return out


# ------------------------------------------------------------------------------------------------
# a4: mel filterbank (construction time, torch ops in the reference's order -> bit-identical matrix)
# ------------------------------------------------------------------------------------------------
def _hertz_to_mel(hz, htk):
    """functional.py:26-45."""
    hz = torch.as_tensor(hz).type(torch.get_default_dtype())
    if htk:
        return 2595. * torch.log10(torch.tensor(1., dtype=torch.get_default_dtype()) + (hz / 700.))
    slope = 200.0 / 3
    knee_hz = 1000.0
    knee_mel = (knee_hz - 0.0) / slope
    logstep = math.log(6.4) / 27.0
    return torch.where(hz >= knee_hz, knee_mel + torch.log(hz / knee_hz) / logstep, (hz - 0.0) / slope)


def _mel_to_hertz(mel, htk):
    """functional.py:5-23."""
    mel = torch.as_tensor(mel).type(torch.get_default_dtype())
    if htk:
        return 700. * (10 ** (mel / 2595.) - 1.)
    slope = 200.0 / 3
    knee_hz = 1000.0
    knee_mel = (knee_hz - 0.0) / slope
    logstep = math.log(6.4) / 27.0
    return torch.where(mel >= knee_mel, knee_hz * torch.exp(logstep * (mel - knee_mel)), 0.0 + slope * mel)


def create_mel_filter(num_freqs, num_mels, min_freq, max_freq, htk):
    """Triangular mel weights `(num_freqs, num_mels)`, Slaney scale unless `htk`, no area
    normalisation (functional.py:131-169).  Runs once per module on the host with the same fp32
    torch operators in the same order as the reference, so the matrix is bit-identical to the
    reference's (tests assert `torch.equal` against golden matrices)."""
    lo_mel = _hertz_to_mel(min_freq, htk)
    hi_mel = _hertz_to_mel(max_freq, htk)
    bin_hz = torch.linspace(min_freq, max_freq, num_freqs)                    # :155
    edge_hz = _mel_to_hertz(torch.linspace(lo_mel, hi_mel, num_mels + 2), htk)   # :158-159
    width = edge_hz[1:] - edge_hz[:-1]                                        # :160
    offset = edge_hz.unsqueeze(0) - bin_hz.unsqueeze(1)                       # :163  (num_freqs, num_mels + 2)
    falling = (-1. * offset[:, :-2]) / width[:-1]                             # :165
    rising = offset[:, 2:] / width[1:]                                        # :166
    return torch.clamp(torch.min(falling, rising), min=0.)                    # :167


# ------------------------------------------------------------------------------------------------
# a3: apply_filterbank (tcgen05)
# ------------------------------------------------------------------------------------------------
class FilterbankPlan(object):
    """Device image of a `(num_freqs, num_bands)` matrix prepared for the tensor-core kernel:
    tf32 hi/lo split, 128B-swizzled K-major operand blocks per 32-bin slice, zero blocks skipped
    (tac_fbplan_build_host).  Built once per matrix per device and cached by the caller."""

    def __init__(self, filterbank, device):
        fb = filterbank.detach().to(device="cpu", dtype=torch.float32).contiguous()
        if fb.dim() != 2:
            raise RuntimeError("apply_filterbank: filterbank must be (num_freqs, num_bands), got %s"
                               % (tuple(fb.shape),))
        self.num_freqs, self.num_bands = int(fb.size(0)), int(fb.size(1))
        lib = _cabi.lib()
        cap = int(lib.tac_fbplan_bytes(self.num_freqs, self.num_bands))
        if cap <= 0:
            raise RuntimeError("apply_filterbank: unsupported filterbank shape %s" % (tuple(fb.shape),))
        host = torch.empty(cap, dtype=torch.uint8)
        used = ctypes.c_int64(0)
        _cabi.check(lib.tac_fbplan_build_host(_cabi.ptr(fb), self.num_freqs, self.num_bands, _cabi.ptr(host), cap,
                                              ctypes.byref(used)))
        # non-zero: the matrix is a chain of two-band rows and the one-kernel path applies (n_fft = 2048)
        self.band_handle = int(lib.tac_fbplan_band_handle(_cabi.ptr(host)))
        # non-zero: the one-kernel mel path applies at the fft length this matrix belongs to (2048: band plan; 256 / 512 /
        # 1024: per-band bin ranges); it is the handle tac_melspec_banded_f32 takes
        self.fft_length = 2 * (self.num_freqs - 1)
        self.fused_handle = int(lib.tac_fbplan_fused_handle(_cabi.ptr(host), self.fft_length))
        self.blob = host[:used.value].to(device)
        self.device = self.blob.device
        # What the plan was built from: address, layout and in-place version counter of the matrix.  The plan keeps the
        # matrix's STORAGE alive, so the address cannot be recycled for another tensor while the plan is cached (the
        # stale-plan hazard of keying by address alone); aliases of the same storage (`.detach()`, as the backward pass
        # passes) share address and version counter and hit the same plan.  Writes through `.data` bypass the version
        # counter, as they do for autograd: `invalidate_filterbank_plans()` / a fresh module cache cover that case.
        self.storage = filterbank.untyped_storage()
        self.key = FilterbankPlan.key_of(filterbank)

    @staticmethod
    def key_of(filterbank):
        return (filterbank.data_ptr(), filterbank._version, tuple(filterbank.shape), tuple(filterbank.stride()), str(filterbank.device))

    def built_from(self, filterbank, device):
        return self.key == FilterbankPlan.key_of(filterbank) and self.device == torch.device(device)


# plans of matrices passed to the functional API (modules keep their own): a handful, evicted oldest first
_GLOBAL_PLANS = collections.OrderedDict()
_GLOBAL_PLANS_MAX = 8
_plans_lock = threading.Lock()


def invalidate_filterbank_plans():
    """Drop every cached plan of the functional API (after editing a filterbank through `.data`)."""
    with _plans_lock:
        _GLOBAL_PLANS.clear()


def _plan_for(filterbank, device, cache=None):
    if cache is not None:
        plan = cache.get("plan")
        if plan is not None and plan.built_from(filterbank, device):
            return plan
        plan = FilterbankPlan(filterbank, device)
        cache["plan"] = plan
        return plan
    slot = FilterbankPlan.key_of(filterbank) + (str(torch.device(device)),)
    with _plans_lock:
        plan = _GLOBAL_PLANS.get(slot)
        if plan is not None and plan.built_from(filterbank, device):
            _GLOBAL_PLANS.move_to_end(slot)
            return plan
    plan = FilterbankPlan(filterbank, device)
    with _plans_lock:
        _GLOBAL_PLANS[slot] = plan
        _GLOBAL_PLANS.move_to_end(slot)
        while len(_GLOBAL_PLANS) > _GLOBAL_PLANS_MAX:
            _GLOBAL_PLANS.popitem(last=False)
    return plan


def _power_mel(spec, is_complex, power, plan, to_db, ref, amin):
    shape = spec.shape[:-1] if is_complex else spec.shape
    if len(shape) < 2:
        raise RuntimeError("apply_filterbank: expected (*, num_freqs, time%s), got %s"
                           % (", 2" if is_complex else "", tuple(spec.shape)))
    n_bins, frames = int(shape[-2]), int(shape[-1])
    if n_bins != plan.num_freqs:
        raise RuntimeError("apply_filterbank: spectrogram has %d frequency bins, filterbank has %d rows"
                           % (n_bins, plan.num_freqs))
    lead = tuple(shape[:-2])
    n_seq = 1
    for d in lead:
        n_seq *= int(d)
    out = torch.empty(lead + (plan.num_bands, frames), dtype=torch.float32, device=spec.device)
    with torch.cuda.device(spec.device):
        _cabi.check(_cabi.lib().tac_power_mel_f32(
            _cabi.ptr(spec), int(is_complex), float(power), n_seq, frames, n_bins, _cabi.ptr(plan.blob),
            plan.num_bands, int(bool(to_db)), float(ref), float(amin), _cabi.ptr(out), _cabi.stream_ptr(spec.device)))
    return out


def apply_filterbank(mag_specgrams, filterbank, _cache=None):
    """`(*, num_freqs, time) x (num_freqs, num_bands) -> (*, num_bands, time)`: contraction over
    the frequency axis (functional.py:172-184) on the tcgen05 tensor cores with 3xTF32 split
    accumulation (csrc/melbank.cu).  Any dense matrix is accepted; zero blocks are skipped."""
    if isinstance(mag_specgrams, torch.Tensor) and mag_specgrams.dtype == torch.float64:
        return _f64.apply_filterbank(mag_specgrams, filterbank)
    if _wants_grad(mag_specgrams) or _wants_grad(filterbank):
        return _ApplyFilterbankFn.apply(mag_specgrams, filterbank, _cache)
    spec = _as_f32_cuda(mag_specgrams, "mag_specgrams")
    plan = _plan_for(filterbank, spec.device, _cache)
    return _power_mel(spec, False, 1.0, plan, False, 1.0, 1e-7)


# ------------------------------------------------------------------------------------------------
# a5: amplitude_to_db
# ------------------------------------------------------------------------------------------------
def amplitude_to_db(x, ref=1.0, amin=1e-7):
    """`10 * (log10(max(x^2, amin)) - log10(ref))` (functional.py:277-296; note the square)."""
    if isinstance(x, torch.Tensor) and x.dtype == torch.float64:
        return _f64.amplitude_to_db(x, float(ref), float(amin))
    if _wants_grad(x):
        return _AmplitudeToDbFn.apply(x, float(ref), float(amin))
    a = _as_f32_cuda(x, "x")
    out = torch.empty_like(a)
    with torch.cuda.device(a.device):
        _cabi.check(_cabi.lib().tac_amplitude_to_db_f32(_cabi.ptr(a), a.numel(), float(ref), float(amin), _cabi.ptr(out),
                                                        _cabi.stream_ptr(a.device)))
    return out


def db_to_amplitude(x, ref=1.0):
    """`sqrt(10 ** (x / 10 + log10(ref)))` (functional.py:299-314): the inverse of `amplitude_to_db`."""
    if isinstance(x, torch.Tensor) and x.dtype == torch.float64:
        return _f64.db_to_amplitude(x, float(ref))
    if _wants_grad(x):
        return _DbToAmplitudeFn.apply(x, float(ref))
    a = _as_f32_cuda(x, "x")
    out = torch.empty_like(a)
    with torch.cuda.device(a.device):
        _cabi.check(_cabi.lib().tac_db_to_amplitude_f32(_cabi.ptr(a), a.numel(), float(ref), _cabi.ptr(out),
                                                        _cabi.stream_ptr(a.device)))
    return out


# ------------------------------------------------------------------------------------------------
# N4: angle / magphase
# ------------------------------------------------------------------------------------------------
def _magphase(complex_tensor, power, want_mag, name):
    if isinstance(complex_tensor, torch.Tensor) and complex_tensor.dtype == torch.float64:
        return _f64.magphase(complex_tensor, float(power), want_mag, name)
    if _wants_grad(complex_tensor):
        mag, phase = _MagphaseFn.apply(complex_tensor, float(power), want_mag, name)
        return (mag if want_mag else None), phase
    z = _as_f32_cuda(complex_tensor, "complex_tensor")
    if z.dim() < 1 or z.size(-1) != 2:
        raise RuntimeError("%s: expected a (*, 2) tensor, got %s" % (name, tuple(z.shape)))
    phase = torch.empty(z.shape[:-1], dtype=torch.float32, device=z.device)
    mag = torch.empty_like(phase) if want_mag else None
    with torch.cuda.device(z.device):
        _cabi.check(_cabi.lib().tac_magphase_f32(_cabi.ptr(z), phase.numel(), float(power),
                                                 _cabi.ptr(mag) if want_mag else None, _cabi.ptr(phase),
                                                 _cabi.stream_ptr(z.device)))
    return mag, phase


def angle(complex_tensor):
    """`atan2(im, re)` of a `(*, 2)` tensor (functional.py:187-191)."""
    return _magphase(complex_tensor, 1.0, False, "angle")[1]


def magphase(complex_tensor, power=1.):
    """`(complex_norm(z, power), angle(z))` (functional.py:194-201), one pass over `z`."""
    return _magphase(complex_tensor, power, True, "magphase")


# ------------------------------------------------------------------------------------------------
# N2: phase vocoder
# ------------------------------------------------------------------------------------------------
_pv_tables = {}


def _phase_vocoder_tables(n_in, rate, device):
    """Index / interpolation tables of functional.py:239-243 and :250-253, built with the reference's own
    expressions in the reference's dtype (torch's default dtype, as `torch.arange(0, T, rate)` has there), on
    the host, once per (T, rate): `time_steps.long()`, `(time_steps + 1).long()`, `remainder(time_steps, 1)`."""
    key = (int(n_in), float(rate), torch.get_default_dtype(), str(device))
    hit = _pv_tables.get(key)
    if hit is None:
        time_steps = torch.arange(0, n_in, rate)
        alphas = torch.remainder(time_steps, torch.tensor(1.))
        hit = (time_steps.long().to(torch.int32).to(device), (time_steps + 1).long().to(torch.int32).to(device),
               alphas.to(torch.float64).to(device))
        if len(_pv_tables) > 64:
            _pv_tables.clear()
        _pv_tables[key] = hit
    return hit


def _phase_vocoder_ranges(idx0, idx1, n_in):
    """(n_in, 2) int32 tables: the range [lo, hi) of output steps whose idx0 / idx1 equals each input frame (both index
    tables are monotone), for the gather of the backward kernel."""
    frames = torch.arange(n_in, device=idx0.device, dtype=torch.int32)
    out = []
    for idx in (idx0, idx1):
        lo = torch.searchsorted(idx, frames, right=False)
        hi = torch.searchsorted(idx, frames, right=True)
        out.append(torch.stack([lo, hi], dim=1).to(torch.int32).contiguous())
    return out


def _phase_vocoder_prepare(complex_specgrams, rate, phase_advance):
    _cabi.require_cuda(complex_specgrams, "complex_specgrams")
    if complex_specgrams.dtype not in (torch.float32, torch.float64):
        raise NotImplementedError("phase_vocoder: dtype %s (float32 and float64 are implemented)" % complex_specgrams.dtype)
    spec = complex_specgrams.contiguous()
    if spec.dim() < 3 or spec.size(-1) != 2:
        raise RuntimeError("phase_vocoder: expected (*, num_freqs, time, 2), got %s" % (tuple(spec.shape),))
    if not rate > 0:
        raise ValueError("phase_vocoder: rate must be positive, got %r" % (rate,))
    n_bins, n_in = int(spec.size(-3)), int(spec.size(-2))
    adv = phase_advance.detach().to(device=spec.device, dtype=spec.dtype).reshape(-1).contiguous()
    if adv.numel() != n_bins:
        raise RuntimeError("phase_vocoder: phase_advance has %d entries for %d frequency bins" % (adv.numel(), n_bins))
    idx0, idx1, alphas = _phase_vocoder_tables(n_in, rate, spec.device)
    lead = tuple(spec.shape[:-3])
    n_seq = 1
    for d in lead:
        n_seq *= int(d)
    return spec, adv, idx0, idx1, alphas, lead, n_seq, n_bins, n_in


def _phase_vocoder_forward(complex_specgrams, rate, phase_advance):
    spec, adv, idx0, idx1, alphas, lead, n_seq, n_bins, n_in = _phase_vocoder_prepare(complex_specgrams, rate, phase_advance)
    n_out = int(idx0.numel())
    out = torch.empty(lead + (n_bins, n_out, 2), dtype=spec.dtype, device=spec.device)
    fn = _cabi.lib().tac_phase_vocoder_f32 if spec.dtype == torch.float32 else _cabi.lib().tac_phase_vocoder_f64
    with torch.cuda.device(spec.device):
        _cabi.check(fn(_cabi.ptr(spec), n_seq, n_bins, n_in, _cabi.ptr(idx0), _cabi.ptr(idx1), _cabi.ptr(alphas),
                       _cabi.ptr(adv), n_out, _cabi.ptr(out), _cabi.stream_ptr(spec.device)))
    return out


class _PhaseVocoderFn(torch.autograd.Function):
    """functional.py:204-274 differentiates through index_select / atan2 / norm / cumsum / cos / sin; here one gather
    kernel (csrc/phase_vocoder.cu: phase_vocoder_backward_kernel), float64 inside like the forward pass."""

    @staticmethod
    def forward(ctx, complex_specgrams, rate, phase_advance):
        ctx.save_for_backward(complex_specgrams)
        ctx.rate, ctx.phase_advance = rate, phase_advance
        return _phase_vocoder_forward(complex_specgrams.detach(), rate, phase_advance)

    @staticmethod
    def backward(ctx, grad_out):
        (z,) = ctx.saved_tensors
        spec, adv, idx0, idx1, alphas, lead, n_seq, n_bins, n_in = _phase_vocoder_prepare(z.detach(), ctx.rate, ctx.phase_advance)
        n_out = int(idx0.numel())
        g = grad_out.to(dtype=spec.dtype).contiguous()
        range0, range1 = _phase_vocoder_ranges(idx0, idx1, n_in)
        ws = torch.empty(max(n_seq * n_bins * n_out * 16, 1), dtype=torch.uint8, device=spec.device)
        grad = torch.empty_like(spec)
        fn = _cabi.lib().tac_phase_vocoder_backward_f32 if spec.dtype == torch.float32 else _cabi.lib().tac_phase_vocoder_backward_f64
        with torch.cuda.device(spec.device):
            _cabi.check(fn(_cabi.ptr(spec), _cabi.ptr(g), n_seq, n_bins, n_in, _cabi.ptr(idx0), _cabi.ptr(idx1), _cabi.ptr(alphas),
                           _cabi.ptr(adv), n_out, _cabi.ptr(range0), _cabi.ptr(range1), _cabi.ptr(ws), ws.numel(), _cabi.ptr(grad),
                           _cabi.stream_ptr(spec.device)))
        return grad, None, None


def phase_vocoder(complex_specgrams, rate, phase_advance):
    """Time-stretch a complex STFT by `rate` without changing pitch (functional.py:204-274).
    `(*, channel, num_freqs, time, 2) -> (*, channel, num_freqs, ceil(time / rate), 2)`; `phase_advance` is the
    `(num_freqs, 1)` expected phase advance per bin.  float32 or float64 tensors; angles, the phase wrap, the
    running phase sum and sin / cos are evaluated in float64 either way (the reference's float32 evaluation
    loses the accumulated phase, which is why its own test runs in float64 -- tests/test_functional.py:85-88),
    so the result matches the reference called in float64 on the same values.  Differentiable w.r.t. the spectrogram
    (one gather kernel, deterministic); `phase_advance` is a constant (it raises if it requires grad)."""
    _no_param_grad(phase_advance, "phase_vocoder: phase_advance")
    if _wants_grad(complex_specgrams):
        return _PhaseVocoderFn.apply(complex_specgrams, rate, phase_advance)
    return _phase_vocoder_forward(complex_specgrams, rate, phase_advance)


# ------------------------------------------------------------------------------------------------
# a6: fused Melspectrogram pipeline
# ------------------------------------------------------------------------------------------------
def _workspace(device, nbytes):
    """Scratch for the power tiles of the two-kernel path, allocated per call: torch's caching allocator recycles it
    stream-safely (a process-wide buffer was shared by concurrent streams / threads on one device)."""
    return torch.empty(max(nbytes, 1), dtype=torch.uint8, device=device)


def _mel_frame_major(filterbank, fft_length, layout, device, cache):
    """True when `melspectrogram` will write a `(*, frames, num_bands)` buffer (one-kernel path, reference layout)."""
    if layout != "reference" or os.environ.get("TAC_MELSPEC_FUSED", "1") == "0":
        return False
    plan = _plan_for(filterbank, device, cache)
    return bool(plan.fused_handle) and plan.fft_length == int(fft_length)


def melspectrogram(waveforms, filterbank, fft_length, hop_length=None, win_length=None, window=None,
                   center=True, pad_mode='reflect', normalized=False, power=2.0,
                   to_db=False, ref=1.0, amin=1e-7, layout="reference", _cache=None, _raw_buffer=False):
    """`Melspectrogram(...)(x)` (layers.py:307-347), optionally with `AmplitudeToDb` appended
    (layers.py:350-381).  `(*, channel, time) -> (*, channel, num_bands, frames)`.

    `fft_length == 2048` and a filterbank whose rows have at most two adjacent non-zeros (every
    triangular filterbank): ONE kernel, two frames per warp from their samples to their `num_bands` outputs,
    the spectrum never leaves the SM (csrc/stft_pair.cu; `TAC_MELSPEC_FUSED=0` disables).  `fft_length` 256 / 512 /
    1024 and a filterbank whose bands cover short bin ranges (again every triangular one): ONE kernel too, the warp
    FFT kernels with the band sums as their epilogue (csrc/stft_multi.cu OUT_MEL_RANGE).
    Otherwise two back-to-back kernels: stft + |.|^power into frame-major power tiles that stay in
    L2, then the tensor-core filterbank with the dB clamp in its epilogue.

    `layout="reference"` (default; one-kernel path only) returns the reference's memory order -- a
    transposed view of a `(*, frames, num_bands)` buffer, exactly the strides of `apply_filterbank`'s
    `matmul(...).transpose(-2, -1)` (functional.py:183-184; a frame's bands are one 512-byte store);
    `layout="contiguous"` returns a contiguous `(*, num_bands, frames)` tensor (4-byte stores, ~9 % slower
    at BASELINE config 2).  The two-kernel path always returns a contiguous tensor."""
    if _wants_grad(window) and not (isinstance(waveforms, torch.Tensor) and waveforms.dtype == torch.float64):
        # a learnable window: the stages, each with its own adjoint kernels; contiguous (*, bands, frames) result
        spec = complex_norm(stft(waveforms, fft_length, hop_length, win_length, window, center, pad_mode, normalized, True), power)
        mel = apply_filterbank(spec, filterbank, _cache=_cache)
        return amplitude_to_db(mel, ref, amin) if to_db else mel
    if isinstance(waveforms, torch.Tensor) and waveforms.dtype == torch.float64:     # the stages, in double (csrc/f64_path.cu)
        _no_param_grad(window, "melspectrogram (float64): window")
        spec = _f64.complex_norm(_f64.stft(waveforms, fft_length, hop_length, win_length, window, center, pad_mode, normalized,
                                           True), float(power))
        mel = _f64.apply_filterbank(spec, filterbank)
        return _f64.amplitude_to_db(mel, float(ref), float(amin)) if to_db else mel
    if _wants_grad(waveforms) or _wants_grad(filterbank):
        # The Function returns the buffer as it lies in memory; the reference-layout view is taken out here, where
        # autograd sees an ordinary transpose (a view created inside Function.forward costs an as_strided replay:
        # measured 0.46 ms per backward call at BASELINE config 2).
        kw = dict(fft_length=fft_length, hop_length=hop_length, win_length=win_length, window=window, center=center,
                  pad_mode=pad_mode, normalized=normalized, power=power, to_db=to_db, ref=ref, amin=amin, layout=layout)
        frame_major = _mel_frame_major(filterbank, fft_length, layout, waveforms.device, _cache)
        buf = _MelspectrogramFn.apply(waveforms, filterbank, kw, _cache, frame_major)
        return buf.transpose(-2, -1) if frame_major else buf
    x = _as_f32_cuda(waveforms, "waveforms")
    hop, lead, flat, frames = _stft_geometry(x, fft_length, hop_length, center)
    win = _frame_window(window, win_length, fft_length, x.device)
    plan = _plan_for(filterbank, x.device, _cache)
    if plan.num_freqs != fft_length // 2 + 1:
        raise RuntimeError("melspectrogram: filterbank has %d rows, stft yields %d bins"
                           % (plan.num_freqs, fft_length // 2 + 1))
    if layout not in ("contiguous", "reference"):
        raise ValueError("melspectrogram: layout must be 'contiguous' or 'reference', got %r" % (layout,))
    lib = _cabi.lib()
    frames = max(frames, 0)
    if plan.fused_handle and plan.fft_length == int(fft_length) and os.environ.get("TAC_MELSPEC_FUSED", "1") != "0":
        frame_major = layout == "reference"
        shape = (flat.size(0), frames, plan.num_bands) if frame_major else (flat.size(0), plan.num_bands, frames)
        out = torch.empty(shape, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _cabi.check(lib.tac_melspec_banded_f32(
                *_stft_args(flat, win, fft_length, hop, center, pad_mode, normalized),
                float(power), _cabi.ptr(plan.blob), plan.fused_handle, plan.num_bands, int(bool(to_db)), float(ref),
                float(amin), _cabi.ptr(out), int(frame_major), _cabi.stream_ptr(x.device)))
        out = out.reshape(lead + out.shape[1:])
        return out if _raw_buffer else (out.transpose(-2, -1) if frame_major else out)
    ws_bytes = int(lib.tac_melspec_workspace_bytes(flat.size(0), flat.size(1), int(fft_length), hop, int(bool(center))))
    ws = _workspace(x.device, ws_bytes)
    out = torch.empty((flat.size(0), plan.num_bands, frames), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _cabi.check(lib.tac_melspec_f32(
            *_stft_args(flat, win, fft_length, hop, center, pad_mode, normalized),
            float(power), _cabi.ptr(plan.blob), plan.num_bands, int(bool(to_db)), float(ref), float(amin),
            _cabi.ptr(ws), ws.numel(), _cabi.ptr(out), _cabi.stream_ptr(x.device)))
    return out.reshape(lead + out.shape[1:])


# ------------------------------------------------------------------------------------------------
# N4 / H7: backward passes (csrc/stft_backward.cu).  The reference is differentiable through torch's
# operators (functional.py:99-107, :126-128, :183-184, :291-296); these Functions give the same
# d loss / d signal with hand-written adjoint kernels.  Forward = the kernels above (grad mode is
# off inside Function.forward, so the public functions take their plain path).
# ------------------------------------------------------------------------------------------------
def _grad_f32(g):
    if g.dtype != torch.float32:
        g = g.float()
    return g


def _stft_backward_call(x_or_none, grad_out, shape, fft_length, hop_length, win_length, window, center, pad_mode,
                        normalized, onesided, power):
    """grad w.r.t. the waveform of `stft` (power None: grad_out (*, bins, frames, 2)) or `spectrogram`."""
    device = grad_out.device
    n_samples = int(shape[-1])
    n_seq = 1
    for d in shape[:-1]:
        n_seq *= int(d)
    hop = fft_length // 4 if hop_length is None else int(hop_length)
    win = _frame_window(window, win_length, fft_length, device)
    g = _grad_f32(grad_out).contiguous()
    gx = torch.empty((n_seq, n_samples), dtype=torch.float32, device=device)
    lib = _cabi.lib()
    ws_bytes = int(lib.tac_stft_backward_workspace_bytes(n_seq, n_samples, int(fft_length), hop, int(bool(center))))
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=device)          # frame gradients before the overlap-add
    with torch.cuda.device(device):
        if power is None:
            _cabi.check(lib.tac_stft_backward_f32(
                _cabi.ptr(g), n_seq, n_samples, _cabi.ptr(win), int(fft_length), hop, int(bool(center)),
                _cabi.PAD_MODES[pad_mode], int(bool(normalized)), int(bool(onesided)), _cabi.ptr(gx), _cabi.ptr(ws), ws_bytes,
                _cabi.stream_ptr(device)))
        else:
            flat = x_or_none.reshape(-1, n_samples)
            _cabi.check(lib.tac_spectrogram_backward_f32(
                *_stft_args(flat, win, fft_length, hop, center, pad_mode, normalized), int(bool(onesided)), float(power),
                _cabi.ptr(g), _cabi.ptr(gx), _cabi.ptr(ws), ws_bytes, _cabi.stream_ptr(device)))
    return gx.reshape(shape)


def _window_grad_call(x, grad_out, shape, fft_length, hop_length, win_length, window, center, pad_mode, normalized, onesided):
    """dL/d window of `stft` (functional.py:93-107 under autograd): the frame gradients before the window multiply
    (tac_stft_backward_f32 run with a window of ones leaves them, times the normalisation, in its workspace), reduced
    against the padded waveform by tac_window_grad_f32; returned in the window's own shape (win_length) and dtype."""
    device = grad_out.device
    n_samples = int(shape[-1])
    n_seq = 1
    for d in shape[:-1]:
        n_seq *= int(d)
    hop = fft_length // 4 if hop_length is None else int(hop_length)
    g = _grad_f32(grad_out).contiguous()
    flat = x.reshape(-1, n_samples)
    lib = _cabi.lib()
    ones = torch.ones(fft_length, dtype=torch.float32, device=device)
    gx = torch.empty((n_seq, n_samples), dtype=torch.float32, device=device)           # by-product, discarded
    ws_bytes = int(lib.tac_stft_backward_workspace_bytes(n_seq, n_samples, int(fft_length), hop, int(bool(center))))
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=device)
    scratch = torch.empty(256 * int(fft_length), dtype=torch.uint8, device=device)
    gw = torch.empty(fft_length, dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        _cabi.check(lib.tac_stft_backward_f32(
            _cabi.ptr(g), n_seq, n_samples, _cabi.ptr(ones), int(fft_length), hop, int(bool(center)),
            _cabi.PAD_MODES[pad_mode], int(bool(normalized)), int(bool(onesided)), _cabi.ptr(gx), _cabi.ptr(ws), ws_bytes,
            _cabi.stream_ptr(device)))
        _cabi.check(lib.tac_window_grad_f32(
            _cabi.ptr(flat), n_seq, n_samples, flat.stride(0) if n_seq > 1 else n_samples, _cabi.ptr(ws), int(fft_length), hop,
            int(bool(center)), _cabi.PAD_MODES[pad_mode], _cabi.ptr(gw), _cabi.ptr(scratch), scratch.numel(), _cabi.stream_ptr(device)))
    wl = fft_length if win_length is None else int(win_length)
    left = (fft_length - wl) // 2                                  # the window sits in the middle of the frame (_frame_window)
    return gw[left:left + wl].to(device=window.device, dtype=window.dtype)


class _StftFn(torch.autograd.Function):
    """`window_param` is the window when it requires grad (a learnable window), else None."""

    @staticmethod
    def forward(ctx, waveforms, window_param, args):
        ctx.args, ctx.shape, ctx.in_dtype = args, tuple(waveforms.shape), waveforms.dtype
        ctx.has_window = window_param is not None
        fft_length, hop_length, win_length, window, center, pad_mode, normalized, onesided = args
        if ctx.has_window:
            ctx.save_for_backward(_as_f32_cuda(waveforms.detach(), "waveforms"))
        ctx.args = (fft_length, hop_length, win_length, window.detach() if isinstance(window, torch.Tensor) else window, center,
                    pad_mode, normalized, onesided)
        return stft(waveforms.detach(), *ctx.args)

    @staticmethod
    def backward(ctx, grad_out):
        fft_length, hop_length, win_length, window, center, pad_mode, normalized, onesided = ctx.args
        gx = gw = None
        if ctx.needs_input_grad[0]:
            gx = _stft_backward_call(None, grad_out, ctx.shape, fft_length, hop_length, win_length, window, center, pad_mode,
                                     normalized, onesided, None).to(ctx.in_dtype)
        if ctx.has_window and ctx.needs_input_grad[1]:
            (x,) = ctx.saved_tensors
            gw = _window_grad_call(x, grad_out, ctx.shape, fft_length, hop_length, win_length, window, center, pad_mode,
                                   normalized, onesided)
        return gx, gw, None


class _SpectrogramFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, waveforms, args):
        x = _as_f32_cuda(waveforms.detach(), "waveforms")
        ctx.save_for_backward(x)
        ctx.args, ctx.shape = args, tuple(waveforms.shape)
        return spectrogram(x, *args)

    @staticmethod
    def backward(ctx, grad_out):
        (x,) = ctx.saved_tensors
        fft_length, hop_length, win_length, window, center, pad_mode, normalized, onesided, power = ctx.args
        return _stft_backward_call(x, grad_out, ctx.shape, fft_length, hop_length, win_length, window, center, pad_mode,
                                   normalized, onesided, power), None


def _complex_norm_backward(z, grad_out, power):
    g = _grad_f32(grad_out).contiguous()
    gz = torch.empty_like(z)
    with torch.cuda.device(z.device):
        _cabi.check(_cabi.lib().tac_complex_norm_backward_f32(_cabi.ptr(z), _cabi.ptr(g), g.numel(), float(power),
                                                              _cabi.ptr(gz), _cabi.stream_ptr(z.device)))
    return gz


class _ComplexNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, complex_tensor, power):
        z = _as_f32_cuda(complex_tensor.detach(), "complex_tensor")
        ctx.save_for_backward(z)
        ctx.power = power
        return complex_norm(z, power)

    @staticmethod
    def backward(ctx, grad_out):
        (z,) = ctx.saved_tensors
        return _complex_norm_backward(z, grad_out, ctx.power), None


def _filterbank_backward(grad_y, filterbank):
    """grad_y: (*, num_bands, frames) with any strides in its last two dims -> (*, num_freqs, frames) contiguous."""
    g = _grad_f32(grad_y)
    lead = tuple(g.shape[:-2])
    n_bands, frames = int(g.size(-2)), int(g.size(-1))
    g3 = g.reshape((-1, n_bands, frames))                 # a view whenever the leading dims are contiguous (the usual case)
    fb = filterbank.detach().to(device=g.device, dtype=torch.float32).contiguous()
    if fb.dim() != 2 or fb.size(1) != n_bands:
        raise RuntimeError("apply_filterbank backward: filterbank %s does not match %d bands" % (tuple(fb.shape), n_bands))
    out = torch.empty((g3.size(0), fb.size(0), frames), dtype=torch.float32, device=g.device)
    with torch.cuda.device(g.device):
        _cabi.check(_cabi.lib().tac_filterbank_backward_f32(
            _cabi.ptr(g3), g3.stride(0) if g3.size(0) > 1 else n_bands * frames, g3.stride(1), g3.stride(2), _cabi.ptr(fb),
            g3.size(0), frames, int(fb.size(0)), n_bands, _cabi.ptr(out), _cabi.stream_ptr(g.device)))
    return out.reshape(lead + (int(fb.size(0)), frames))


def _filterbank_param_grad(spec, grad_y, like):
    """d loss / d filterbank[k, m] = sum over sequences and frames of spec[.., k, t] * grad_y[.., m, t]: a plain
    (num_freqs x frames) . (frames x num_bands) GEMM per sequence, summed -- a library GEMM (torch.matmul), off the
    hot path; the reference gets the same gradient from its matmul (functional.py:183)."""
    g = _grad_f32(grad_y)
    s3 = spec.reshape((-1,) + tuple(spec.shape[-2:]))
    g3 = g.reshape((-1,) + tuple(g.shape[-2:]))
    tf32 = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        gfb = torch.matmul(s3, g3.transpose(-2, -1)).sum(0)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = tf32
    return gfb.to(dtype=like.dtype, device=like.device)


class _ApplyFilterbankFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mag_specgrams, filterbank, cache):
        ctx.filterbank = filterbank
        spec = _as_f32_cuda(mag_specgrams.detach(), "mag_specgrams")
        ctx.save_for_backward(spec if filterbank.requires_grad else spec.new_empty(0))
        return apply_filterbank(spec, filterbank.detach(), cache)

    @staticmethod
    def backward(ctx, grad_out):
        (spec,) = ctx.saved_tensors
        gspec = _filterbank_backward(grad_out, ctx.filterbank) if ctx.needs_input_grad[0] else None
        gfb = _filterbank_param_grad(spec, grad_out, ctx.filterbank) if ctx.needs_input_grad[1] else None
        return gspec, gfb, None


def _amplitude_to_db_backward(x, grad_out, amin):
    g = _grad_f32(grad_out).contiguous()
    gx = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _cabi.check(_cabi.lib().tac_amplitude_to_db_backward_f32(_cabi.ptr(x), _cabi.ptr(g), g.numel(), float(amin),
                                                                 _cabi.ptr(gx), _cabi.stream_ptr(x.device)))
    return gx


class _AmplitudeToDbFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, ref, amin):
        a = _as_f32_cuda(x.detach(), "x")
        ctx.save_for_backward(a)
        ctx.amin = amin
        return amplitude_to_db(a, ref, amin)

    @staticmethod
    def backward(ctx, grad_out):
        (a,) = ctx.saved_tensors
        return _amplitude_to_db_backward(a, grad_out, ctx.amin), None, None


class _MelspectrogramFn(torch.autograd.Function):
    """Backward of the fused chain: [dB adjoint on the recomputed mel values ->] filterbank adjoint ->
    spectrogram adjoint (which recomputes the spectrum from the saved waveform).  Nothing but the waveform is
    saved by the forward pass."""

    @staticmethod
    def forward(ctx, waveforms, filterbank, kw, cache, frame_major):
        x = _as_f32_cuda(waveforms.detach(), "waveforms")
        ctx.save_for_backward(x)
        ctx.filterbank, ctx.kw, ctx.cache, ctx.shape, ctx.frame_major = filterbank, kw, cache, tuple(waveforms.shape), frame_major
        return melspectrogram(x, filterbank.detach(), _cache=cache, _raw_buffer=True, **kw)     # the buffer, not the view

    @staticmethod
    def backward(ctx, grad_out):
        (x,) = ctx.saved_tensors
        kw = ctx.kw
        g = grad_out.transpose(-2, -1) if ctx.frame_major else grad_out                # logical (*, bands, frames)
        if kw["to_db"]:
            plain = dict(kw, to_db=False, layout="contiguous")
            mel = melspectrogram(x, ctx.filterbank.detach(), _cache=ctx.cache, **plain)  # (*, bands, frames) contiguous
            g = _amplitude_to_db_backward(mel, g, kw["amin"])
        gfb = None
        if ctx.needs_input_grad[1]:                          # learnable filterbank: |X|^p recomputed, one GEMM
            spec = spectrogram(x, kw["fft_length"], kw["hop_length"], kw["win_length"], kw["window"], kw["center"],
                               kw["pad_mode"], kw["normalized"], True, kw["power"])
            gfb = _filterbank_param_grad(spec, g, ctx.filterbank)
            del spec
        if not ctx.needs_input_grad[0]:
            return None, gfb, None, None, None
        # filterbank adjoint -> stft + |.|^p adjoint -> overlap-add in one C-ABI call (tac_melspec_backward_f32)
        g = _grad_f32(g)
        n_bands, frames = int(g.size(-2)), int(g.size(-1))
        g3 = g.reshape((-1, n_bands, frames))
        fb = ctx.filterbank.detach().to(device=x.device, dtype=torch.float32).contiguous()
        fft_length = int(kw["fft_length"])
        hop = fft_length // 4 if kw["hop_length"] is None else int(kw["hop_length"])
        flat = x.reshape(-1, x.size(-1))
        win = _frame_window(kw["window"], kw["win_length"], fft_length, x.device)
        lib = _cabi.lib()
        ws_bytes = int(lib.tac_melspec_backward_workspace_bytes(flat.size(0), flat.size(1), fft_length, hop, int(bool(kw["center"]))))
        ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=x.device)
        gx = torch.empty_like(flat)
        with torch.cuda.device(x.device):
            _cabi.check(lib.tac_melspec_backward_f32(
                *_stft_args(flat, win, fft_length, hop, kw["center"], kw["pad_mode"], kw["normalized"]), float(kw["power"]),
                _cabi.ptr(fb), n_bands, _cabi.ptr(g3), g3.stride(0) if g3.size(0) > 1 else n_bands * frames, g3.stride(1),
                g3.stride(2), _cabi.ptr(gx), _cabi.ptr(ws), ws_bytes, _cabi.stream_ptr(x.device)))
        return gx.reshape(ctx.shape), gfb, None, None, None


def _pointwise_backward(op, a, g1, g2, n, p0, out):
    with torch.cuda.device(a.device):
        _cabi.check(_cabi.lib().tac_pointwise_backward_f32(op, _cabi.ptr(a), _cabi.ptr(g1) if g1 is not None else None,
                                                           _cabi.ptr(g2) if g2 is not None else None, n, float(p0),
                                                           _cabi.ptr(out), _cabi.stream_ptr(a.device)))
    return out


class _DbToAmplitudeFn(torch.autograd.Function):
    """functional.py:299-314 differentiates through torch.pow; here dy/dx = y ln(10) / 20 from the saved output."""

    @staticmethod
    def forward(ctx, x, ref):
        y = db_to_amplitude(x.detach(), ref)
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, grad_out):
        (y,) = ctx.saved_tensors
        g = _grad_f32(grad_out).contiguous()
        return _pointwise_backward(0, y, g, None, y.numel(), 0.0, torch.empty_like(y)), None


class _MagphaseFn(torch.autograd.Function):
    """magphase / angle (functional.py:187-201): both outputs' gradients fold into one dL/dz pass."""

    @staticmethod
    def forward(ctx, complex_tensor, power, want_mag, name):
        z = _as_f32_cuda(complex_tensor.detach(), "complex_tensor")
        ctx.save_for_backward(z)
        ctx.power, ctx.want_mag = power, want_mag
        mag, phase = _magphase(z, power, want_mag, name)
        if not want_mag:
            mag = phase.new_empty(0)
            ctx.mark_non_differentiable(mag)
        return mag, phase

    @staticmethod
    def backward(ctx, grad_mag, grad_phase):
        (z,) = ctx.saved_tensors
        gm = _grad_f32(grad_mag).contiguous() if (ctx.want_mag and grad_mag is not None) else None
        gp = _grad_f32(grad_phase).contiguous() if grad_phase is not None else None
        if gm is None and gp is None:
            return torch.zeros_like(z), None, None, None
        return _pointwise_backward(1, z, gm, gp, z.numel() // 2, ctx.power, torch.empty_like(z)), None, None, None


class _MuLawDecodingFn(torch.autograd.Function):
    """Float codes are differentiable in the reference (functional.py:349-354, no rounding on that path)."""

    @staticmethod
    def forward(ctx, x_mu, n_quantize):
        c = _as_f32_cuda(x_mu.detach(), "x_mu")
        ctx.save_for_backward(c)
        ctx.mu = float(n_quantize - 1)
        return mu_law_decoding(c, n_quantize)

    @staticmethod
    def backward(ctx, grad_out):
        (c,) = ctx.saved_tensors
        g = _grad_f32(grad_out).contiguous()
        return _pointwise_backward(2, c, g, None, c.numel(), ctx.mu, torch.empty_like(c)), None


class PreparedMelspectrogram(object):
    """The `melspectrogram` call with everything shape-independent resolved once (window on the device,
    filterbank plan, argument marshalling): `prepared(x, out)` is then a single C-ABI call into a
    caller-owned output -- what a serving loop or a CUDA-graph capture wants.  `x`: contiguous float32
    CUDA tensor `(*, channel, time)` of the shape given at construction; `out`: a buffer from
    `empty_output()`.  Returns the `(*, channel, num_bands, frames)` result: `out` itself, or with
    `layout="reference"` on the one-kernel path the transposed view of the `(*, frames, num_bands)` buffer
    (the reference's memory order, see `melspectrogram`)."""

    def __init__(self, shape, device, filterbank, fft_length, hop_length=None, win_length=None, window=None,
                 center=True, pad_mode='reflect', normalized=False, power=2.0, to_db=False, ref=1.0, amin=1e-7,
                 layout="reference"):
        self.device = torch.device(device)
        if self.device.type == "cuda" and self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.shape = tuple(int(d) for d in shape)
        self.n_samples = self.shape[-1]
        self.n_seq = 1
        for d in self.shape[:-1]:
            self.n_seq *= d
        lib = _cabi.lib()
        self.hop = fft_length // 4 if hop_length is None else int(hop_length)
        self.frames = max(int(lib.tac_stft_num_frames(self.n_samples, int(fft_length), self.hop, int(bool(center)))), 0)
        self.window = _frame_window(window, win_length, fft_length, self.device)
        self.plan = _plan_for(filterbank, self.device, None)
        if self.plan.num_freqs != fft_length // 2 + 1:
            raise RuntimeError("melspectrogram: filterbank has %d rows, stft yields %d bins"
                               % (self.plan.num_freqs, fft_length // 2 + 1))
        if pad_mode not in _cabi.PAD_MODES:
            raise NotImplementedError("stft: pad_mode=%r" % (pad_mode,))
        self.fused = (bool(self.plan.fused_handle) and self.plan.fft_length == int(fft_length)
                      and os.environ.get("TAC_MELSPEC_FUSED", "1") != "0")
        self.fft_length = int(fft_length)
        if layout not in ("contiguous", "reference"):
            raise ValueError("layout must be 'contiguous' or 'reference', got %r" % (layout,))
        self.frame_major = layout == "reference" and self.fused     # `out` is then (*, frames, num_bands); view it transposed
        self.out_shape = self.shape[:-1] + ((self.frames, self.plan.num_bands) if self.frame_major
                                            else (self.plan.num_bands, self.frames))
        head = [self.n_seq, self.n_samples, self.n_samples, _cabi.ptr(self.window), int(fft_length), self.hop,
                int(bool(center)), _cabi.PAD_MODES[pad_mode], int(bool(normalized)), float(power), _cabi.ptr(self.plan.blob)]
        tail = [self.plan.num_bands, int(bool(to_db)), float(ref), float(amin)]
        if self.fused:
            self._fn, self._head, self._tail = lib.tac_melspec_banded_f32, head + [self.plan.fused_handle] + tail, [int(self.frame_major)]
            self._ws = None
        else:
            nbytes = int(lib.tac_melspec_workspace_bytes(self.n_seq, self.n_samples, int(fft_length), self.hop, int(bool(center))))
            self._ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=self.device)
            self._fn, self._head, self._tail = lib.tac_melspec_f32, head + tail + [_cabi.ptr(self._ws), self._ws.numel()], []

    def empty_output(self):
        return torch.empty(self.out_shape, dtype=torch.float32, device=self.device)

    def __call__(self, x, out):
        if (tuple(x.shape) != self.shape or x.dtype != torch.float32 or not x.is_contiguous() or x.device != self.device
                or tuple(out.shape) != self.out_shape or out.dtype != torch.float32 or not out.is_contiguous()
                or out.device != self.device):
            raise RuntimeError("PreparedMelspectrogram: expected contiguous float32 x %s and out %s on %s"
                               % (self.shape, self.out_shape, self.device))
        with torch.cuda.device(self.device):
            _cabi.check(self._fn(_cabi.ptr(x), *self._head, _cabi.ptr(out), *self._tail,
                                 ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))
        return out.transpose(-2, -1) if self.frame_major else out

    def gather_into(self, x, gathered, item_offset=None):
        """Batch-sharded call (BASELINE config 4): compute this rank's `x` and store the result into the full
        output of EVERY rank (`gathered`: a `distributed.PeerGatheredOutput` or `MulticastGatheredOutput` of shape
        `(total_items,) + out_shape[1:]`) from the kernel's epilogue, over NVLink -- the all-gather of SURVEY 8(e)
        without a collective.  `item_offset`: this rank's first batch item in the full output (default:
        `shard_range` of equal shards).  Follow with `gathered.barrier()`.  One-kernel path only."""
        if not self.fused or self.fft_length != 2048:
            raise NotImplementedError("gather_into: only the one-kernel mel path (fft_length 2048, triangular filterbank) "
                                      "stores to peer buffers")
        if tuple(x.shape) != self.shape or x.dtype != torch.float32 or not x.is_contiguous() or x.device != self.device:
            raise RuntimeError("PreparedMelspectrogram: expected contiguous float32 x %s on %s" % (self.shape, self.device))
        if len(self.shape) < 2:
            raise RuntimeError("gather_into: x needs a leading batch dimension to shard")
        full = tuple(gathered.tensor.shape)
        if len(full) != len(self.out_shape) or full[1:] != self.out_shape[1:] or gathered.device != self.device:
            raise RuntimeError("gather_into: gathered buffer %s does not match (*,) + %s" % (full, self.out_shape[1:]))
        if item_offset is None:
            item_offset = gathered.rank * self.shape[0]
        if item_offset < 0 or item_offset + self.shape[0] > full[0]:
            raise RuntimeError("gather_into: items [%d, %d) outside the gathered batch of %d"
                               % (item_offset, item_offset + self.shape[0], full[0]))
        seq_per_item = self.n_seq // max(self.shape[0], 1)
        if hasattr(gathered, "mc_payload"):                    # distributed.MulticastGatheredOutput: one store per value
            with torch.cuda.device(self.device):
                _cabi.check(_cabi.lib().tac_melspec_banded_mc_f32(
                    _cabi.ptr(x), *self._head, ctypes.c_void_p(gathered.mc_payload),
                    int(item_offset) * seq_per_item, int(self.frame_major),
                    ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))
            return gathered.tensor.transpose(-2, -1) if self.frame_major else gathered.tensor
        with torch.cuda.device(self.device):
            _cabi.check(_cabi.lib().tac_melspec_banded_peers_f32(
                _cabi.ptr(x), *self._head, ctypes.cast(gathered.payload_array, ctypes.c_void_p), gathered.world,
                int(item_offset) * seq_per_item, int(self.frame_major),
                ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))
        return gathered.tensor.transpose(-2, -1) if self.frame_major else gathered.tensor


# ------------------------------------------------------------------------------------------------
# a7 / a8: mu-law
# ------------------------------------------------------------------------------------------------
def mu_law_encoding(x, n_quantize=256):
    """mu-law companding to int64 codes, no clamping of the input (functional.py:317-335).
    Bit-exact with the reference's fp32 CPU chain for every finite float (see _mulaw_tables)."""
    _forward_only(x, "mu_law_encoding")
    _cabi.require_cuda(x, "x")
    if x.dtype == torch.float64:                           # the reference's formula in the input's dtype (functional.py:331-334)
        return _f64.mu_law_encoding(x, int(n_quantize))
    if not x.dtype.is_floating_point:
        x = x.to(torch.float)                              # functional.py:329-330
    a = _as_f32_cuda(x, "x")
    thr, idx_min, x_limit = _mulaw_tables.on_device("enc", n_quantize, a.device)
    out = torch.empty(a.shape, dtype=torch.int64, device=a.device)
    with torch.cuda.device(a.device):
        _cabi.check(_cabi.lib().tac_mulaw_encode_f32_i64(
            _cabi.ptr(a), a.numel(), int(n_quantize), _cabi.ptr(thr), thr.numel(), int(idx_min), float(x_limit),
            _cabi.ptr(out), _cabi.stream_ptr(a.device)))
    return out


def mu_law_decoding(x_mu, n_quantize=256, dtype=torch.float32):
    """mu-law expansion of codes to float32 (functional.py:338-354)."""
    _cabi.require_cuda(x_mu, "x_mu")
    if dtype == torch.float64 and not x_mu.dtype.is_floating_point:
        return _f64.mu_law_decoding(x_mu, int(n_quantize))
    if dtype != torch.float32 and not x_mu.dtype.is_floating_point:
        raise NotImplementedError("mu_law_decoding: float32 and float64 outputs are implemented, not %s" % dtype)
    (lut,) = _mulaw_tables.on_device("dec", n_quantize, x_mu.device)
    lib = _cabi.lib()
    if x_mu.dtype.is_floating_point:
        if _wants_grad(x_mu):
            return _MuLawDecodingFn.apply(x_mu, int(n_quantize))
        codes = _as_f32_cuda(x_mu, "x_mu")
        fn = lib.tac_mulaw_decode_f32_f32
    else:
        codes = x_mu.to(torch.int64).contiguous()
        fn = lib.tac_mulaw_decode_i64_f32
    out = torch.empty(codes.shape, dtype=torch.float32, device=codes.device)
    with torch.cuda.device(codes.device):
        _cabi.check(fn(_cabi.ptr(codes), codes.numel(), int(n_quantize), _cabi.ptr(lut), _cabi.ptr(out),
                       _cabi.stream_ptr(codes.device)))
    return out
