"""Parity oracle: the reference's torch fp32 CPU path, restated.  TEST INFRASTRUCTURE ONLY.

Every function names the reference lines it restates (paths relative to the reference
checkout, commit bdbbb18).  The operators, their order and their dtypes are kept identical
so that results are bit-identical to the reference run on the same host; the code itself is
an independent restatement (checked by ``oracle/gen_golden.py`` with ``torch.equal``).

The reference calls ``torch.stft`` in its pre-1.8 form (real ``(..., 2)`` output).  Torch
2.x only offers the complex form, so the restatement asks for ``return_complex=True`` and
views the result as real -- the same bytes the legacy call produced.
"""
import math

import torch

__all__ = [
    "stft", "complex_norm", "hertz_to_mel", "mel_to_hertz", "create_mel_filter",
    "apply_filterbank", "amplitude_to_db", "mu_law_encoding", "mu_law_decoding",
    "spectrogram", "melspectrogram", "mel_filterbank_for",
    "angle", "magphase", "db_to_amplitude", "phase_vocoder", "hpss",
]

_SLANEY_HZ_PER_MEL = 200.0 / 3          # linear region slope   (functional.py:15,37)
_SLANEY_KNEE_HZ = 1000.0                # start of log region   (functional.py:18,42)
_SLANEY_LOGSTEP = math.log(6.4) / 27.0  # log region step       (functional.py:20,44)


def hertz_to_mel(hz, htk):
    """functional.py:26-45 (`_hertz_to_mel`)."""
    hz = torch.as_tensor(hz).type(torch.get_default_dtype())
    if htk:
        one = torch.tensor(1., dtype=torch.get_default_dtype())
        return 2595. * torch.log10(one + (hz / 700.))
    linear = (hz - 0.0) / _SLANEY_HZ_PER_MEL
    knee_mel = (_SLANEY_KNEE_HZ - 0.0) / _SLANEY_HZ_PER_MEL
    logpart = knee_mel + torch.log(hz / _SLANEY_KNEE_HZ) / _SLANEY_LOGSTEP
    return torch.where(hz >= _SLANEY_KNEE_HZ, logpart, linear)


def mel_to_hertz(mel, htk):
    """functional.py:5-23 (`_mel_to_hertz`)."""
    mel = torch.as_tensor(mel).type(torch.get_default_dtype())
    if htk:
        return 700. * (10 ** (mel / 2595.) - 1.)
    linear = 0.0 + _SLANEY_HZ_PER_MEL * mel
    knee_mel = (_SLANEY_KNEE_HZ - 0.0) / _SLANEY_HZ_PER_MEL
    logpart = _SLANEY_KNEE_HZ * torch.exp(_SLANEY_LOGSTEP * (mel - knee_mel))
    return torch.where(mel >= knee_mel, logpart, linear)


def create_mel_filter(num_freqs, num_mels, min_freq, max_freq, htk):
    """functional.py:131-169: triangular mel weights, shape (num_freqs, num_mels), no area norm."""
    mel_lo = hertz_to_mel(min_freq, htk)
    mel_hi = hertz_to_mel(max_freq, htk)
    bin_hz = torch.linspace(min_freq, max_freq, num_freqs)          # :155
    edges_hz = mel_to_hertz(torch.linspace(mel_lo, mel_hi, num_mels + 2), htk)   # :158-159
    widths = edges_hz[1:] - edges_hz[:-1]                           # :160
    dist = edges_hz.unsqueeze(0) - bin_hz.unsqueeze(1)              # :163
    falling = (-1. * dist[:, :-2]) / widths[:-1]                    # :165
    rising = dist[:, 2:] / widths[1:]                               # :166
    return torch.clamp(torch.min(falling, rising), min=0.)          # :167


def mel_filterbank_for(num_mels=128, sample_rate=22050, min_freq=0.0, max_freq=None,
                       fft_length=None, htk=False):
    """Filterbank exactly as `Melspectrogram` builds it: layers.py:330-344 with
    `MelFilterbank.__init__` (layers.py:184-195): num_freqs from fft_length, max_freq
    defaults to the *integer* sample_rate // 2."""
    num_freqs = fft_length // 2 + 1 if fft_length else 1025
    if sample_rate is None and max_freq is None:
        raise ValueError("need max_freq or sample_rate")
    top = max_freq if max_freq else sample_rate // 2
    return create_mel_filter(num_freqs, num_mels, min_freq, top, htk)


def stft(waveforms, fft_length, hop_length=None, win_length=None, window=None,
         center=True, pad_mode='reflect', normalized=False, onesided=True):
    """functional.py:48-113.  (*, channel, time) -> (*, channel, freq, frames, 2)."""
    lead = waveforms.shape[:-1]
    flat = waveforms.reshape(-1, waveforms.size(-1))                # :89-91
    if window is None:                                              # :93-97
        window = torch.hann_window(fft_length if win_length is None else win_length)
    spec = torch.stft(flat, n_fft=fft_length, hop_length=hop_length, win_length=win_length,
                      window=window, center=center, pad_mode=pad_mode,
                      normalized=normalized, onesided=onesided, return_complex=True)
    spec = torch.view_as_real(spec)                                 # legacy (..., 2) layout
    return spec.reshape(lead + spec.shape[1:])                      # :109-111


def complex_norm(complex_tensor, power=1.0):
    """functional.py:116-128: L2 norm over the last (re, im) axis, then `.pow(power)`."""
    mag = torch.norm(complex_tensor, 2, -1)
    return mag if power == 1.0 else mag.pow(power)


def apply_filterbank(mag_specgrams, filterbank):
    """functional.py:172-184: contraction over the freq axis, output (..., bands, time)."""
    return torch.matmul(mag_specgrams.transpose(-2, -1), filterbank).transpose(-2, -1)


def amplitude_to_db(x, ref=1.0, amin=1e-7):
    """functional.py:277-296.  Squares its input first (quirk 1 of SURVEY section 0)."""
    sq = torch.clamp(x.pow(2.), min=amin)
    ref_t = torch.tensor(ref, device=x.device, requires_grad=False, dtype=x.dtype)
    return 10.0 * (torch.log10(sq) - torch.log10(ref_t))


def mu_law_encoding(x, n_quantize=256):
    """functional.py:317-335: companding then truncation to int64; no clamp of the input."""
    if not x.dtype.is_floating_point:
        x = x.to(torch.float)
    mu = torch.tensor(n_quantize - 1, dtype=x.dtype, requires_grad=False)
    comp = x.sign() * torch.log1p(mu * x.abs()) / torch.log1p(mu)   # :333
    return ((comp + 1) / 2 * mu + 0.5).long()                        # :334


def mu_law_decoding(x_mu, n_quantize=256, dtype=torch.float32):
    """functional.py:338-354 (default dtype is frozen to float32 at import, :338)."""
    if not x_mu.dtype.is_floating_point:
        x_mu = x_mu.to(dtype)
    mu = torch.tensor(n_quantize - 1, dtype=x_mu.dtype, requires_grad=False)
    y = (x_mu / mu) * 2 - 1.                                        # :352
    return y.sign() * (torch.exp(y.abs() * torch.log1p(mu)) - 1.) / mu   # :353


def spectrogram(x, fft_length, hop_length=None, win_length=None, window=None, center=True,
                pad_mode='reflect', normalized=False, onesided=True, power=1.):
    """`Spectrogram(...)(x)`: layers.py:267-304 = STFT then ComplexNorm(power).
    The layer path always has a window buffer (layers.py:76-82)."""
    if window is None:
        window = torch.hann_window(fft_length if win_length is None else win_length)
    z = stft(x, fft_length, hop_length, win_length, window, center, pad_mode, normalized, onesided)
    return complex_norm(z, power)


def melspectrogram(x, num_mels=128, sample_rate=22050, min_freq=0.0, max_freq=None, htk=False,
                   to_db=False, ref=1.0, amin=1e-7, **stft_kwargs):
    """`Melspectrogram(...)(x)` (layers.py:307-347), optionally followed by
    `AmplitudeToDb(ref, amin)` (layers.py:350-381) as in BASELINE config 3."""
    fb = mel_filterbank_for(num_mels, sample_rate, min_freq, max_freq,
                            stft_kwargs.get('fft_length', None), htk)
    p = spectrogram(x, power=2., **stft_kwargs)
    mel = apply_filterbank(p, fb)
    return amplitude_to_db(mel, ref, amin) if to_db else mel


def angle(complex_tensor):
    """functional.py:187-191."""
    return torch.atan2(complex_tensor[..., 1], complex_tensor[..., 0])


def magphase(complex_tensor, power=1.):
    """functional.py:194-201."""
    return complex_norm(complex_tensor, power), angle(complex_tensor)


def db_to_amplitude(x, ref=1.0):
    """functional.py:299-314: 10 ** (x / 10 + log10(ref)), then the square root."""
    ref_t = torch.tensor(ref, device=x.device, requires_grad=False, dtype=x.dtype)
    return torch.pow(10.0, x / 10.0 + torch.log10(ref_t)).pow(0.5)


def phase_vocoder(complex_specgrams, rate, phase_advance):
    """functional.py:204-274, dtype-generic like the reference (its value test runs it in float64 with
    float64 as torch's default dtype, tests/test_functional.py:76-93).
    (*, num_freqs, time, 2) -> (*, num_freqs, ceil(time / rate), 2)."""
    spec = complex_specgrams
    lead = [slice(None)] * (spec.dim() - 2)                                   # :236-237
    steps = torch.arange(0, spec.size(-2), rate, device=spec.device)          # :239-240 (default dtype)
    frac = torch.remainder(steps, torch.tensor(1., device=spec.device))       # :242-243
    first_phase = angle(spec[tuple(lead + [slice(1)])])                              # :244
    padded = torch.nn.functional.pad(spec, [0, 0, 0, 2])                      # :247-248
    lo = padded[tuple(lead + [steps.long()])]                                        # :250-251
    hi = padded[tuple(lead + [(steps + 1).long()])]                                  # :253-254
    ang_lo, ang_hi = angle(lo), angle(hi)                                     # :256-257
    mag_lo, mag_hi = torch.norm(lo, dim=-1), torch.norm(hi, dim=-1)           # :259-260
    dphi = ang_hi - ang_lo - phase_advance                                    # :262
    dphi = dphi - 2 * math.pi * torch.round(dphi / (2 * math.pi))             # :263
    dphi = dphi + phase_advance                                               # :266
    dphi = torch.cat([first_phase, dphi[tuple(lead + [slice(-1)])]], dim=-1)         # :267
    running = torch.cumsum(dphi, -1)                                          # :268
    mag = frac * mag_hi + (1 - frac) * mag_lo                                 # :270
    return torch.stack([mag * torch.cos(running), mag * torch.sin(running)], dim=-1)   # :272-279


def hpss(mag_specgrams, kernel_size=31, power=2.0, hard=False, mask_only=False):
    """beta_hpss.py:37-129 (a beta module outside the reference's package namespace): median filtering along frequency
    (percussive) and time (harmonic) of the reflect-padded magnitudes, `^power`, soft (eps 1e-6) or hard masks.
    (batch, ch, freq, time) -> (harmonic, percussive, mask_harm, mask_perc)."""
    if isinstance(kernel_size, int):
        kernel_size = (kernel_size, kernel_size)                             # :100-101
    k_perc, k_harm = kernel_size
    pads = (k_perc // 2, k_perc // 2, k_harm // 2, k_harm // 2)               # :103-104 (time by the first size, freq by the second)
    padded = torch.nn.functional.pad(mag_specgrams, pad=pads, mode='reflect')  # :107
    perc = torch.empty_like(mag_specgrams)
    harm = torch.empty_like(mag_specgrams)
    off_t, off_f = k_harm // 2, k_perc // 2
    for f in range(perc.shape[2]):                                            # :88-90: window over frequency, time un-padded
        perc[:, :, f, :] = torch.median(padded[:, :, f:f + k_perc, off_t:-off_t], dim=2)[0]
    for t in range(harm.shape[3]):                                            # :84-86: window over time, frequency un-padded
        harm[:, :, :, t] = torch.median(padded[:, :, off_f:-off_f, t:t + k_harm], dim=3)[0]
    if power != 1.0:                                                          # :94-95
        perc.pow_(power)
        harm.pow_(power)
    eps = 1e-6                                                                # :97
    if hard:                                                                  # :116-118
        mask_harm, mask_perc = harm > perc, harm < perc
    else:                                                                     # :120-121
        mask_harm = (harm + eps) / (harm + perc + eps)
        mask_perc = (perc + eps) / (harm + perc + eps)
    if mask_only:                                                             # :123-124
        return None, None, mask_harm, mask_perc
    return mag_specgrams * mask_harm, mag_specgrams * mask_perc, mask_harm, mask_perc   # :126
