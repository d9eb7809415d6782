"""float64 operator path (csrc/f64_path.cu).  The reference is dtype-generic -- functional.py:99 hands a double waveform to
torch.stft and every later stage computes in the dtype it is given -- so double tensors are served by plain double
kernels with the same operator surface.  Forward only: a double input that requires grad raises.  The tuned float32
kernels are the product path; this one exists so a double-precision check of a pipeline runs on the same package.
"""
import ctypes

import torch

from . import _cabi

_F64 = torch.float64


def _check(t, name):
    _cabi.require_cuda(t, name)
    if torch.is_grad_enabled() and t.requires_grad:
        raise NotImplementedError("%s: float64 tensors are forward-only here (the adjoint kernels are float32)" % name)
    return t.contiguous()


def _window(window, win_length, fft_length, device):
    if win_length is None:
        win_length = fft_length
    if window is None:
        window = torch.hann_window(win_length)               # functional.py:93-97: float32 values, promoted by torch.stft
    if window.dim() != 1 or window.size(0) != win_length:
        raise RuntimeError("stft: expected a 1-D window of size win_length=%d, got %s" % (win_length, tuple(window.shape)))
    if not (0 < win_length <= fft_length):
        raise RuntimeError("stft: expected 0 < win_length <= n_fft, got win_length=%d n_fft=%d" % (win_length, fft_length))
    window = window.to(device=device, dtype=_F64)
    if win_length < fft_length:
        left = (fft_length - win_length) // 2
        window = torch.nn.functional.pad(window, (left, fft_length - win_length - left))
    return window.contiguous()


def stft(waveforms, fft_length, hop_length, win_length, window, center, pad_mode, normalized, onesided):
    x = _check(waveforms, "waveforms")
    if pad_mode not in _cabi.PAD_MODES:
        raise NotImplementedError("stft: pad_mode=%r (supported: %s)" % (pad_mode, sorted(_cabi.PAD_MODES)))
    hop = fft_length // 4 if hop_length is None else int(hop_length)
    lead, n_samples = x.shape[:-1], x.size(-1)
    flat = x.reshape(-1, n_samples)
    lib = _cabi.lib()
    frames = max(int(lib.tac_stft_num_frames(n_samples, fft_length, hop, int(bool(center)))), 0)
    win = _window(window, win_length, fft_length, x.device)
    bins = fft_length // 2 + 1 if onesided else fft_length
    out = torch.empty((flat.size(0), bins, frames, 2), dtype=_F64, device=x.device)
    with torch.cuda.device(x.device):
        _cabi.check(lib.tac_stft_f64(
            _cabi.ptr(flat), flat.size(0), n_samples, flat.stride(0) if flat.size(0) > 1 else n_samples, _cabi.ptr(win),
            int(fft_length), hop, int(bool(center)), _cabi.PAD_MODES[pad_mode], int(bool(normalized)), int(bool(onesided)),
            _cabi.ptr(out), _cabi.stream_ptr(x.device)))
    return out.reshape(lead + out.shape[1:])


def complex_norm(z, power):
    z = _check(z, "complex_tensor")
    if z.dim() < 1 or z.size(-1) != 2:
        raise RuntimeError("complex_norm: expected a (*, 2) tensor, got %s" % (tuple(z.shape),))
    out = torch.empty(z.shape[:-1], dtype=_F64, device=z.device)
    with torch.cuda.device(z.device):
        _cabi.check(_cabi.lib().tac_complex_norm_f64(_cabi.ptr(z), out.numel(), ctypes.c_double(power), _cabi.ptr(out),
                                                     _cabi.stream_ptr(z.device)))
    return out


def magphase(z, power, want_mag, name):
    z = _check(z, "complex_tensor")
    if z.dim() < 1 or z.size(-1) != 2:
        raise RuntimeError("%s: expected a (*, 2) tensor, got %s" % (name, tuple(z.shape)))
    phase = torch.empty(z.shape[:-1], dtype=_F64, device=z.device)
    mag = torch.empty_like(phase) if want_mag else None
    with torch.cuda.device(z.device):
        _cabi.check(_cabi.lib().tac_magphase_f64(_cabi.ptr(z), phase.numel(), ctypes.c_double(power),
                                                 _cabi.ptr(mag) if want_mag else None, _cabi.ptr(phase), _cabi.stream_ptr(z.device)))
    return mag, phase


def apply_filterbank(spec, filterbank):
    spec = _check(spec, "mag_specgrams")
    if filterbank.dtype != _F64:                               # what torch.matmul says in the reference (functional.py:183)
        raise RuntimeError("apply_filterbank: expected scalar type Double but found %s (the spectrogram is float64)"
                           % str(filterbank.dtype).replace("torch.", "").capitalize())
    if filterbank.dim() != 2:
        raise RuntimeError("apply_filterbank: filterbank must be (num_freqs, num_bands), got %s" % (tuple(filterbank.shape),))
    if spec.dim() < 2 or spec.size(-2) != filterbank.size(0):
        raise RuntimeError("apply_filterbank: spectrogram %s does not match a filterbank with %d rows"
                           % (tuple(spec.shape), filterbank.size(0)))
    fb = _check(filterbank.to(spec.device), "filterbank")
    lead, n_bins, frames = spec.shape[:-2], spec.size(-2), spec.size(-1)
    n_seq = 1
    for d in lead:
        n_seq *= int(d)
    out = torch.empty(tuple(lead) + (fb.size(1), frames), dtype=_F64, device=spec.device)
    with torch.cuda.device(spec.device):
        _cabi.check(_cabi.lib().tac_apply_filterbank_f64(_cabi.ptr(spec), _cabi.ptr(fb), n_seq, frames, n_bins, fb.size(1),
                                                         _cabi.ptr(out), _cabi.stream_ptr(spec.device)))
    return out


def amplitude_to_db(x, ref, amin):
    a = _check(x, "x")
    out = torch.empty_like(a)
    with torch.cuda.device(a.device):
        _cabi.check(_cabi.lib().tac_amplitude_to_db_f64(_cabi.ptr(a), a.numel(), ctypes.c_double(ref), ctypes.c_double(amin),
                                                        _cabi.ptr(out), _cabi.stream_ptr(a.device)))
    return out


def db_to_amplitude(x, ref):
    a = _check(x, "x")
    out = torch.empty_like(a)
    with torch.cuda.device(a.device):
        _cabi.check(_cabi.lib().tac_db_to_amplitude_f64(_cabi.ptr(a), a.numel(), ctypes.c_double(ref), _cabi.ptr(out),
                                                        _cabi.stream_ptr(a.device)))
    return out


_dec_luts = {}


def mu_law_decoding(codes, n_quantize):
    """int codes -> float64 (functional.py:345-353 evaluated in double on the host for the n_quantize table values)."""
    _cabi.require_cuda(codes, "x_mu")
    key = (int(n_quantize), str(codes.device))
    lut = _dec_luts.get(key)
    if lut is None:
        mu = torch.tensor(n_quantize - 1, dtype=_F64)
        y = (torch.arange(n_quantize, dtype=_F64) / mu) * 2 - 1.
        lut = (y.sign() * (torch.exp(y.abs() * torch.log1p(mu)) - 1.) / mu).to(codes.device)
        _dec_luts[key] = lut
    c = codes.to(torch.int64).contiguous()
    out = torch.empty(c.shape, dtype=_F64, device=c.device)
    with torch.cuda.device(c.device):
        _cabi.check(_cabi.lib().tac_mulaw_decode_i64_f64(_cabi.ptr(c), c.numel(), int(n_quantize), _cabi.ptr(lut), _cabi.ptr(out),
                                                         _cabi.stream_ptr(c.device)))
    return out


def mu_law_encoding(x, n_quantize):
    a = _check(x, "x")
    out = torch.empty(a.shape, dtype=torch.int64, device=a.device)
    with torch.cuda.device(a.device):
        _cabi.check(_cabi.lib().tac_mulaw_encode_f64_i64(_cabi.ptr(a), a.numel(), int(n_quantize), _cabi.ptr(out),
                                                         _cabi.stream_ptr(a.device)))
    return out
