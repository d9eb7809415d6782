"""Float64 numpy evaluation of the same maths.  TEST INFRASTRUCTURE ONLY.

Stands in for librosa in the reference's value tests (``tests/test_functional.py:55-66``
uses ``librosa.stft``; ``tests/test_layers.py:65-83`` uses ``librosa.power_to_db``), which
is not installable here.  ``librosa.stft(y, n_fft, hop_length, pad_mode)`` is: centre-pad by
``n_fft // 2``, periodic Hann window, frame at ``hop_length``, ``rfft`` -- restated below in
float64.  SURVEY section 4 records that the reference agrees with exactly this restatement
to 7.8e-6 max-abs at config 1.
"""
import numpy as np

_PAD_MODES = {'reflect': 'reflect', 'constant': 'constant', 'replicate': 'edge', 'circular': 'wrap'}


def hann_periodic(n):
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n, dtype=np.float64) / n)


def centre_pad_window(window, n_fft):
    """torch.stft zero-pads a short window on both sides so it sits in the middle of the frame."""
    window = np.asarray(window, dtype=np.float64)
    left = (n_fft - window.shape[0]) // 2
    out = np.zeros(n_fft, dtype=np.float64)
    out[left:left + window.shape[0]] = window
    return out


def num_frames(n_samples, n_fft, hop, center=True):
    padded = n_samples + (2 * (n_fft // 2) if center else 0)
    return 1 + (padded - n_fft) // hop


def stft(x, n_fft, hop=None, win_length=None, window=None, center=True, pad_mode='reflect',
         normalized=False, onesided=True):
    """x: (..., time) array -> complex128 (..., bins, frames)."""
    x = np.asarray(x, dtype=np.float64)
    hop = n_fft // 4 if hop is None else hop
    if window is None:
        window = hann_periodic(n_fft if win_length is None else win_length)
    w = centre_pad_window(window, n_fft)
    lead = x.shape[:-1]
    flat = x.reshape(-1, x.shape[-1])
    if center:
        flat = np.pad(flat, ((0, 0), (n_fft // 2, n_fft // 2)), mode=_PAD_MODES[pad_mode])
    frames = 1 + (flat.shape[-1] - n_fft) // hop
    idx = np.arange(n_fft)[None, :] + hop * np.arange(frames)[:, None]
    seg = flat[:, idx] * w                                  # (n, frames, n_fft)
    spec = np.fft.rfft(seg, axis=-1) if onesided else np.fft.fft(seg, axis=-1)
    if normalized:
        spec = spec / np.sqrt(n_fft)
    spec = np.swapaxes(spec, -1, -2)                        # (n, bins, frames)
    return spec.reshape(lead + spec.shape[1:])


def power_to_db(s, ref=1.0, amin=1e-7):
    """librosa.power_to_db(S, ref, amin, top_db=None)."""
    return 10.0 * np.log10(np.maximum(s, amin)) - 10.0 * np.log10(ref)


def melspectrogram(x, fb, n_fft, hop, power=2.0, **kw):
    """|stft|^power contracted with a (bins, bands) matrix -> (..., bands, frames)."""
    p = np.abs(stft(x, n_fft, hop, **kw)) ** power
    return np.einsum('...ft,fm->...mt', p, np.asarray(fb, dtype=np.float64))


def phase_vocoder(d, rate, hop):
    """librosa.phase_vocoder(D, rate, hop_length) as published with librosa 0.6 / 0.7 (the oracle of the
    reference's tests/test_functional.py:100-112; librosa itself is not installable here): a sequential walk
    over the output steps with a running phase -- an independent formulation of what the reference computes
    with gathers and a cumsum.  d: complex128 (bins, frames) -> complex128 (bins, ceil(frames / rate))."""
    d = np.asarray(d, dtype=np.complex128)
    bins, frames = d.shape
    steps = np.arange(0, frames, rate, dtype=np.float64)
    out = np.zeros((bins, len(steps)), dtype=np.complex128)
    expected = np.linspace(0, np.pi * hop, bins)
    running = np.angle(d[:, 0])
    d = np.pad(d, [(0, 0), (0, 2)], mode='constant')
    for t, step in enumerate(steps):
        cols = d[:, int(step):int(step + 2)]
        frac = np.mod(step, 1.0)
        mag = (1.0 - frac) * np.abs(cols[:, 0]) + frac * np.abs(cols[:, 1])
        out[:, t] = mag * np.exp(1.0j * running)
        dphase = np.angle(cols[:, 1]) - np.angle(cols[:, 0]) - expected
        dphase = dphase - 2.0 * np.pi * np.round(dphase / (2.0 * np.pi))
        running = running + expected + dphase
    return out
