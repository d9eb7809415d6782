"""Module surface of torchaudio_contrib (reference: torchaudio_contrib/layers.py) over the B200
kernels: same class names, constructor signatures, defaults, buffers, `__repr__` strings and
exceptions; `Spectrogram` / `Melspectrogram` still return an iterable `nn.Sequential` of
`STFT, ComplexNorm[, ApplyFilterbank]`, but of the `FusedSequential` flavour whose forward runs
the fused kernels when its children form a known chain.
"""
import math
import os

import torch
import torch.nn as nn

from . import functional as F

__all__ = [
    "STFT", "ComplexNorm", "ApplyFilterbank", "Filterbank", "MelFilterbank", "Spectrogram",
    "Melspectrogram", "AmplitudeToDb", "MuLawEncoding", "MuLawDecoding", "FusedSequential", "Sequential",
    "TimeStretch", "DbToAmplitude",
]


class _ModuleNoStateBuffers(nn.Module):
    """Buffers (window, filterbank) are rebuilt from constructor arguments, never checkpointed:
    they are dropped from `state_dict()` and ignored on load (reference layers.py:11-32), so
    `Melspectrogram(...).state_dict() == {}` and a strict load of `{}` succeeds."""

    def state_dict(self, *args, **kwargs):
        prefix = kwargs.get("prefix", args[1] if len(args) > 1 else "")
        full = super(_ModuleNoStateBuffers, self).state_dict(*args, **kwargs)
        for name in self._buffers:
            full.pop(prefix + name, None)
        return full

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        hidden, self._buffers = self._buffers, {}
        try:
            return super(_ModuleNoStateBuffers, self)._load_from_state_dict(state_dict, prefix, *args, **kwargs)
        finally:
            self._buffers = hidden


class STFT(_ModuleNoStateBuffers):
    """Short-time Fourier transform layer (reference layers.py:35-109).

    Args are those of the reference / `torch.stft`: `fft_length`, `hop_length` (default
    `fft_length // 4`), `win_length` (default `fft_length`), `window` (default Hann of
    `win_length`), `center`, `pad_mode`, `normalized`, `onesided`.
    Input `(*, channel, time)`, output `(*, channel, num_freqs, frames, 2)`.
    """

    def __init__(self, fft_length, hop_length=None, win_length=None, window=None, center=True,
                 pad_mode='reflect', normalized=False, onesided=True):
        super(STFT, self).__init__()
        self.fft_length = fft_length
        self.hop_length = hop_length
        self.win_length = win_length
        self.center = center
        self.pad_mode = pad_mode
        self.normalized = normalized
        self.onesided = onesided
        if window is None:
            window = torch.hann_window(fft_length if win_length is None else win_length)
        self.register_buffer('window', window)

    def stft_kwargs(self):
        return dict(hop_length=self.hop_length, win_length=self.win_length, window=self.window,
                    center=self.center, pad_mode=self.pad_mode, normalized=self.normalized)

    def forward(self, waveforms):
        return F.stft(waveforms, self.fft_length, onesided=self.onesided, **self.stft_kwargs())

    def __repr__(self):
        return (self.__class__.__name__
                + '(fft_length={}, hop_length={}, win_length={})'.format(self.fft_length, self.hop_length, self.win_length)
                + '(center={}, pad_mode={}, normalized={}, onesided={})'.format(
                    self.center, self.pad_mode, self.normalized, self.onesided))


class ComplexNorm(nn.Module):
    """`(*, 2) -> (*)`, `|z| ** power` (reference layers.py:112-135)."""

    def __init__(self, power=1.0):
        super(ComplexNorm, self).__init__()
        self.power = power

    def forward(self, complex_tensor):
        return F.complex_norm(complex_tensor, self.power)

    def __repr__(self):
        return self.__class__.__name__ + '(power={})'.format(self.power)


class ApplyFilterbank(_ModuleNoStateBuffers):
    """Contracts the frequency axis with a `(num_freqs, num_bands)` matrix held as a buffer
    (reference layers.py:138-155).  The tensor-core operand image of the matrix is derived
    lazily per device and re-derived if the buffer changes."""

    def __init__(self, filterbank):
        super(ApplyFilterbank, self).__init__()
        self.register_buffer('filterbank', filterbank)
        self._plan_cache = {}

    def forward(self, mag_specgrams):
        return F.apply_filterbank(mag_specgrams, self.filterbank, _cache=self._plan_cache)


class Filterbank(object):
    """Base class of filterbank providers (reference layers.py:158-167)."""

    def __init__(self):
        super(Filterbank, self).__init__()

    def get_filterbank(self):
        raise NotImplementedError


class MelFilterbank(Filterbank):
    """Provider of the triangular mel matrix (reference layers.py:170-212).  `max_freq`
    defaults to the integer `sample_rate // 2`."""

    def __init__(self, num_freqs=1025, num_mels=128, min_freq=0.0, max_freq=None, sample_rate=None, htk=False):
        super(MelFilterbank, self).__init__()
        if sample_rate is None and max_freq is None:
            raise ValueError('Either max_freq or sample_rate should be specified.'
                             ', but both are None.')
        self.num_freqs = num_freqs
        self.num_mels = num_mels
        self.min_freq = min_freq
        self.max_freq = max_freq if max_freq else sample_rate // 2
        self.htk = htk

    def get_filterbank(self):
        return F.create_mel_filter(num_freqs=self.num_freqs, num_mels=self.num_mels, min_freq=self.min_freq,
                                   max_freq=self.max_freq, htk=self.htk)

    def __repr__(self):
        # spelling of the reference kept verbatim (layers.py:205-212), including "snum_mels"
        return (self.__class__.__name__ + '(num_freqs={}, snum_mels={}'.format(self.num_freqs, self.num_mels)
                + ', min_freq={}, max_freq={})'.format(self.min_freq, self.max_freq) + ', htk={}'.format(self.htk))


class AmplitudeToDb(_ModuleNoStateBuffers):
    """`10 * (log10(max(x^2, amin)) - log10(ref))` (reference layers.py:350-381)."""

    def __init__(self, ref=1.0, amin=1e-7):
        super(AmplitudeToDb, self).__init__()
        self.ref = ref
        self.amin = amin
        assert ref > amin, "Reference value is expected to be bigger than amin, but I have" \
                           "ref:{} and amin:{}".format(ref, amin)

    def forward(self, x):
        return F.amplitude_to_db(x, ref=self.ref, amin=self.amin)

    def __repr__(self):
        return self.__class__.__name__ + '(ref={}, amin={})'.format(self.ref, self.amin)


class DbToAmplitude(_ModuleNoStateBuffers):
    """`sqrt(10 ** (x / 10 + log10(ref)))` (reference layers.py:384-412)."""

    def __init__(self, ref=1.0):
        super(DbToAmplitude, self).__init__()
        self.ref = ref

    def forward(self, x):
        return F.db_to_amplitude(x, ref=self.ref)

    def __repr__(self):
        return self.__class__.__name__ + '(ref={})'.format(self.ref)


class TimeStretch(_ModuleNoStateBuffers):
    """Stretch a complex STFT in time by `rate` without changing pitch (reference layers.py:215-264).
    `phase_advance = linspace(0, pi * hop_length, num_freqs)[..., None]` is a buffer; `forward(spec,
    overriding_rate=None)` uses `fixed_rate` unless a rate is passed, returns the input itself for rate 1.0 and
    raises `ValueError` when neither rate is given."""

    def __init__(self, hop_length, num_freqs, fixed_rate=None):
        super(TimeStretch, self).__init__()
        self.fixed_rate = fixed_rate
        phase_advance = torch.linspace(0, math.pi * hop_length, num_freqs)[..., None]
        self.register_buffer('phase_advance', phase_advance)

    def forward(self, complex_specgrams, overriding_rate=None):
        if overriding_rate is None:
            rate = self.fixed_rate
            if rate is None:
                raise ValueError("If no fixed_rate is specified"
                                 ", must pass a valid rate to the forward method.")
        else:
            rate = overriding_rate
        if rate == 1.0:
            return complex_specgrams
        return F.phase_vocoder(complex_specgrams, rate, self.phase_advance)

    def __repr__(self):
        return self.__class__.__name__ + '(fixed_rate={})'.format(self.fixed_rate)


class MuLawEncoding(_ModuleNoStateBuffers):
    """mu-law companding to int64 codes (reference layers.py:415-440)."""

    def __init__(self, n_quantize=256):
        super(MuLawEncoding, self).__init__()
        self.n_quantize = n_quantize

    def forward(self, x):
        return F.mu_law_encoding(x, self.n_quantize)

    def __repr__(self):
        return self.__class__.__name__ + '(n_quantize={})'.format(self.n_quantize)


class MuLawDecoding(_ModuleNoStateBuffers):
    """mu-law expansion of codes to float32 (reference layers.py:443-467)."""

    def __init__(self, n_quantize=256):
        super(MuLawDecoding, self).__init__()
        self.n_quantize = n_quantize

    def forward(self, x_mu):
        return F.mu_law_decoding(x_mu, self.n_quantize)

    def __repr__(self):
        return self.__class__.__name__ + '(n_quantize={})'.format(self.n_quantize)


class FusedSequential(nn.Sequential):
    """`nn.Sequential` whose forward recognises the hot chains and runs them fused:

        STFT, ComplexNorm                                   -> one stft+|.|^p kernel
        STFT, ComplexNorm, ApplyFilterbank[, AmplitudeToDb] -> stft+|.|^p rows (L2) -> tcgen05 filterbank[+dB]
        ComplexNorm, ApplyFilterbank[, AmplitudeToDb]       -> |.|^p + tcgen05 filterbank[+dB] in one kernel
        ApplyFilterbank, AmplitudeToDb                      -> filterbank with dB epilogue

    Children stay ordinary modules, so `*Melspectrogram(...)` unpacking (reference layers.py:346,
    tests/test_layers.py:69) and indexing keep working; any other composition runs child by child.
    Use it in place of `nn.Sequential` when appending `AmplitudeToDb` to keep the epilogue fused.
    """

    def forward(self, x):
        mods = list(self)
        i = 0
        while i < len(mods):
            step = self._fused_step(mods, i, x)
            if step is None:
                x = mods[i](x)
                i += 1
            else:
                x, i = step
        return x

    @staticmethod
    def _prepared_mel(st, power, fb, db, x):
        """The steady-state mel call of a module chain: everything shape-independent (window on the device, filterbank plan,
        argument marshalling) is resolved once per (shape, device, window, matrix, options) and kept on the filterbank
        module, so a repeated call is one C-ABI call (the host side of the generic path costs about as much as the
        config-2 kernel takes).  None when the call is not the plain float32 CUDA forward (autograd, other dtypes, CPU
        tensors, strided inputs, the two-kernel path): the generic function then applies its own checks."""
        if (not isinstance(x, torch.Tensor) or not x.is_cuda or x.dtype != torch.float32 or not x.is_contiguous() or x.dim() < 1
                or (torch.is_grad_enabled() and (x.requires_grad or fb.filterbank.requires_grad
                                                 or (isinstance(st.window, torch.Tensor) and st.window.requires_grad)))):
            return None
        win = st.window
        key = (tuple(x.shape), x.device, F.FilterbankPlan.key_of(fb.filterbank),
               (win.data_ptr(), win._version, tuple(win.shape)) if isinstance(win, torch.Tensor) else None,
               st.fft_length, st.hop_length, st.win_length, st.center, st.pad_mode, st.normalized, float(power),
               (float(db.ref), float(db.amin)) if db is not None else None, os.environ.get("TAC_MELSPEC_FUSED", "1"))
        cache = fb._plan_cache
        hit = cache.get("prepared")
        if hit is None or hit[0] != key:
            try:
                prep = F.PreparedMelspectrogram(x.shape, x.device, fb.filterbank, st.fft_length, st.hop_length, st.win_length,
                                                win, st.center, st.pad_mode, st.normalized, power, db is not None,
                                                db.ref if db is not None else 1.0, db.amin if db is not None else 1e-7)
            except Exception:
                return None                                   # let the generic path raise its own, reference-shaped error
            if not prep.fused:
                cache["prepared"] = (key, None)
                return None
            cache["prepared"] = hit = (key, prep)
        prep = hit[1]
        if prep is None:
            return None
        return prep(x, prep.empty_output())

    @staticmethod
    def _fused_step(mods, i, x):
        def kind(j, cls):
            return j < len(mods) and type(mods[j]) is cls

        if kind(i, STFT) and kind(i + 1, ComplexNorm) and mods[i].onesided:
            st, power = mods[i], mods[i + 1].power
            if kind(i + 2, ApplyFilterbank):
                fb = mods[i + 2]
                db = mods[i + 3] if kind(i + 3, AmplitudeToDb) else None
                y = FusedSequential._prepared_mel(st, power, fb, db, x)
                if y is not None:
                    return y, i + (4 if db is not None else 3)
                y = F.melspectrogram(x, fb.filterbank, st.fft_length, power=power,
                                     to_db=db is not None, ref=db.ref if db else 1.0, amin=db.amin if db else 1e-7,
                                     _cache=fb._plan_cache, **st.stft_kwargs())
                return y, i + (4 if db is not None else 3)
            return F.spectrogram(x, st.fft_length, onesided=True, power=power, **st.stft_kwargs()), i + 2
        if F._wants_grad(x):
            return None                              # child by child: every child has its own backward kernel
        if kind(i, ComplexNorm) and kind(i + 1, ApplyFilterbank):
            fb = mods[i + 1]
            db = mods[i + 2] if kind(i + 2, AmplitudeToDb) else None
            F._forward_only(x, "FusedSequential")
            z = F._as_f32_cuda(x, "complex_tensor")
            plan = F._plan_for(fb.filterbank, z.device, fb._plan_cache)
            y = F._power_mel(z, True, mods[i].power, plan, db is not None, db.ref if db else 1.0, db.amin if db else 1e-7)
            return y, i + (3 if db is not None else 2)
        if kind(i, ApplyFilterbank) and kind(i + 1, AmplitudeToDb):
            fb, db = mods[i], mods[i + 1]
            F._forward_only(x, "FusedSequential")
            s = F._as_f32_cuda(x, "mag_specgrams")
            plan = F._plan_for(fb.filterbank, s.device, fb._plan_cache)
            return F._power_mel(s, False, 1.0, plan, True, db.ref, db.amin), i + 2
        return None


Sequential = FusedSequential


def Spectrogram(fft_length, hop_length=None, win_length=None, window=None, center=True, pad_mode='reflect',
                normalized=False, onesided=True, power=1.):
    """Sequential of `[STFT(), ComplexNorm(power)]` (reference layers.py:267-304)."""
    return FusedSequential(
        STFT(fft_length, hop_length, win_length, window, center, pad_mode, normalized, onesided),
        ComplexNorm(power))


def Melspectrogram(num_mels=128, sample_rate=22050, min_freq=0.0, max_freq=None, num_freqs=None, htk=False,
                   mel_filterbank=None, **kwargs):
    """Sequential of `[STFT(), ComplexNorm(2.), ApplyFilterbank(mel matrix)]` (reference
    layers.py:307-347).  As in the reference the `num_freqs` argument is ignored: the number of
    rows of the mel matrix comes from `kwargs['fft_length']` (1025 when absent), and `**kwargs`
    go to `Spectrogram`.  `mel_filterbank` may be a `Filterbank` subclass to build the matrix."""
    fft_length = kwargs.get('fft_length', None)
    num_freqs = fft_length // 2 + 1 if fft_length else 1025
    provider = MelFilterbank if mel_filterbank is None else mel_filterbank
    matrix = provider(num_mels=num_mels, sample_rate=sample_rate, min_freq=min_freq, max_freq=max_freq,
                      num_freqs=num_freqs, htk=htk).get_filterbank()
    return FusedSequential(*Spectrogram(power=2., **kwargs), ApplyFilterbank(matrix))
