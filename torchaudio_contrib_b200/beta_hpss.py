"""Harmonic-percussive source separation (reference: torchaudio_contrib/beta_hpss.py, a beta module the reference keeps out
of its package namespace; same here: `from torchaudio_contrib_b200.beta_hpss import HPSS, hpss`).

Same names, arguments, return tuple and quirks as the reference (`beta_hpss.py:37-129`): median filters of `kernel_size`
along frequency (percussive) and time (harmonic) on the reflect-padded magnitudes, `^power`, soft masks with eps 1e-6 or
hard (boolean) masks, `(mag * mask_harm, mag * mask_perc, mask_harm, mask_perc)` or `(None, None, masks)` with `mask_only`.
One CUDA kernel (csrc/hpss.cu); there is no CPU path.
"""
import torch
import torch.nn as nn

from . import _cabi

__all__ = ["HPSS", "hpss"]


def hpss(mag_specgrams, kernel_size=31, power=2.0, hard=False, mask_only=False):
    """`(batch, ch, freq, time)` magnitudes (not dB) -> `(harmonic, percussive, mask_harm, mask_perc)`
    (beta_hpss.py:37-129).  `kernel_size`: odd int (or a tuple of two equal-half sizes; the reference's slicing only works
    when both paddings agree, beta_hpss.py:106-114, and raises otherwise -- so does this)."""
    if not (isinstance(kernel_size, tuple) or isinstance(kernel_size, int)):
        raise TypeError('kernel_size is expected to be either tuple of input, but it is: %s' % type(kernel_size))
    if isinstance(kernel_size, int):
        kernel_size = (kernel_size, kernel_size)
    if kernel_size[0] // 2 != kernel_size[1] // 2 or kernel_size[0] // 2 == 0:
        # beta_hpss.py:109-114 slices the time axis by the OTHER kernel's half width: with different halves (or a zero one,
        # where `offset:-offset` is empty) the assignment into `out` fails with a shape error in the reference
        raise RuntimeError("hpss: the expanded size of the tensor must match the existing size (kernel halves %d and %d)"
                           % (kernel_size[0] // 2, kernel_size[1] // 2))
    if torch.is_grad_enabled() and isinstance(mag_specgrams, torch.Tensor) and mag_specgrams.requires_grad:
        raise RuntimeError("hpss: forward only (the reference writes its medians in place, beta_hpss.py:84-90)")
    _cabi.require_cuda(mag_specgrams, "mag_specgrams")
    if mag_specgrams.dtype != torch.float32:
        raise NotImplementedError("hpss: float32 magnitudes, got %s" % mag_specgrams.dtype)
    if mag_specgrams.dim() != 4:
        raise RuntimeError("hpss: expected (batch, ch, freq, time), got %s" % (tuple(mag_specgrams.shape),))
    x = mag_specgrams.contiguous()
    k = int(kernel_size[0])
    if kernel_size[0] % 2 == 0 or kernel_size[1] % 2 == 0:  # even windows: torch.median takes the lower middle of k values -- not built
        raise NotImplementedError("hpss: odd kernel sizes are implemented (the reference documents odd sizes)")
    n_seq, n_freq, n_time = x.size(0) * x.size(1), x.size(2), x.size(3)
    mask_h, mask_p = torch.empty_like(x), torch.empty_like(x)
    out_h = out_p = None
    if not mask_only:
        out_h, out_p = torch.empty_like(x), torch.empty_like(x)
    with torch.cuda.device(x.device):
        _cabi.check(_cabi.lib().tac_hpss_f32(
            _cabi.ptr(x), n_seq, n_freq, n_time, int(k), float(power), int(bool(hard)), int(bool(mask_only)),
            _cabi.ptr(out_h) if out_h is not None else None, _cabi.ptr(out_p) if out_p is not None else None,
            _cabi.ptr(mask_h), _cabi.ptr(mask_p), _cabi.stream_ptr(x.device)))
    if hard:
        mask_h, mask_p = mask_h != 0, mask_p != 0           # boolean masks, as `harm > perc` gives (beta_hpss.py:116-118)
    return out_h, out_p, mask_h, mask_p


class HPSS(nn.Module):
    """Wrap `hpss` (beta_hpss.py:13-34)."""

    def __init__(self, kernel_size=31, power=2.0, hard=False, mask_only=False):
        super(HPSS, self).__init__()
        self.kernel_size = kernel_size
        self.power = power
        self.hard = hard
        self.mask_only = mask_only

    def forward(self, mag_specgrams):
        return hpss(mag_specgrams, self.kernel_size, self.power, self.hard, self.mask_only)

    def __repr__(self):
        return self.__class__.__name__ + '(kernel_size={}, power={}, hard={}, mask_only={})'.format(
            self.kernel_size, self.power, self.hard, self.mask_only)
