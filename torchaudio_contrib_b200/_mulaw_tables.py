"""Constant tables for the mu-law kernels (csrc/mulaw.cu).

The reference defines the mu-law quantiser by an fp32 formula (functional.py:329-334) whose
rounding depends on the host's `log1p` implementation.  To be bit-exact with it for every float,
the encoder kernel does not re-derive that transcendental: it uses the quantiser's *decision
levels* -- for each code k the smallest float x with code(x) >= k -- found here once per
`n_quantize` by bisection over the float line, and the decoder uses the table of the
`n_quantize` decoded values (functional.py:349-353).  Both are construction-time constants, in
the same way the Hann window and the mel matrix are; the per-sample work is all on the GPU.

The quantiser is monotone non-decreasing in x (tests sweep all 2^32 floats for n_quantize=256),
which is what makes the decision levels a complete description.
"""
import threading

import torch

_INT64_MIN = -(1 << 63)
_cache = {}
_lock = threading.Lock()


def _quantise(x, n_quantize):
    """The reference's companding + truncation, functional.py:331-334, on CPU fp32."""
    mu = torch.tensor(n_quantize - 1, dtype=torch.float32)
    comp = x.sign() * torch.log1p(mu * x.abs()) / torch.log1p(mu)
    return ((comp + 1) / 2 * mu + 0.5).long()


def _expand(codes, n_quantize):
    """functional.py:351-353 on CPU fp32."""
    mu = torch.tensor(n_quantize - 1, dtype=torch.float32)
    y = (codes / mu) * 2 - 1.
    return y.sign() * (torch.exp(y.abs() * torch.log1p(mu)) - 1.) / mu


def _key_to_float(keys):
    """Monotone int64 key in [0, 2^32) (ordered like the floats it denotes) -> float32 tensor.
    Non-negative floats: key = bits + 2^31.  Negative floats: key = 2^32 - 1 - bits."""
    bits = torch.where(keys >= (1 << 31), keys - (1 << 31), (1 << 32) - 1 - keys)   # unsigned bit pattern
    bits = torch.where(bits >= (1 << 31), bits - (1 << 32), bits)                   # as signed int32
    return bits.to(torch.int32).view(torch.float32)


def _float_to_key(x):
    bits = x.view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    return torch.where(bits >= (1 << 31), (1 << 32) - 1 - bits, bits + (1 << 31))


def encode_tables(n_quantize):
    """-> (thresholds fp32 [n_thr], idx_min, x_limit).

    thresholds[j] = smallest float whose code is >= idx_min + j; thresholds[0] = -inf.
    x_limit = largest |x| for which the reference still produces a finite intermediate
    (beyond it mu*|x| overflows and the int64 conversion yields INT64_MIN).
    """
    key = ("enc", int(n_quantize))
    with _lock:
        if key in _cache:
            return _cache[key]
    with torch.no_grad():
        # largest finite-result magnitude: bisection on "code != INT64_MIN" over positive floats
        lo = torch.tensor([1 << 31], dtype=torch.int64)                       # key of +0.0
        hi = _float_to_key(torch.tensor([3.4028234663852886e38], dtype=torch.float32))
        if int(_quantise(_key_to_float(hi), n_quantize)) != _INT64_MIN:
            lim_key = hi
        else:
            while int(hi - lo) > 1:
                mid = (lo + hi) // 2
                ok = int(_quantise(_key_to_float(mid), n_quantize)) != _INT64_MIN
                lo, hi = (mid, hi) if ok else (lo, mid)
            lim_key = lo
        x_limit = float(_key_to_float(lim_key))
        neg_key = _float_to_key(torch.tensor([-x_limit], dtype=torch.float32))
        idx_min = int(_quantise(torch.tensor([-x_limit]), n_quantize))
        idx_max = int(_quantise(torch.tensor([x_limit]), n_quantize))
        targets = torch.arange(idx_min + 1, idx_max + 1, dtype=torch.int64)   # codes with a finite level
        lo = neg_key.expand_as(targets).clone()       # code(lo) = idx_min < target
        hi = lim_key.expand_as(targets).clone()       # code(hi) = idx_max >= target
        for _ in range(34):
            mid = (lo + hi) // 2
            ge = _quantise(_key_to_float(mid), n_quantize) >= targets
            hi = torch.where(ge, mid, hi)
            lo = torch.where(ge, lo, mid)
        levels = _key_to_float(hi)
        thresholds = torch.cat([torch.tensor([float("-inf")]), levels]).contiguous()
    out = (thresholds, idx_min, x_limit)
    with _lock:
        _cache[key] = out
    return out


def decode_table(n_quantize):
    """-> fp32 [n_quantize]: the reference's decoded value of every in-range code."""
    key = ("dec", int(n_quantize))
    with _lock:
        if key in _cache:
            return _cache[key]
    with torch.no_grad():
        lut = _expand(torch.arange(n_quantize, dtype=torch.float32), n_quantize).contiguous()
    with _lock:
        _cache[key] = lut
    return lut


_device_cache = {}


def on_device(kind, n_quantize, device):
    """Device-resident copy of a table, cached per (kind, n_quantize, device)."""
    k = (kind, int(n_quantize), str(device))
    hit = _device_cache.get(k)
    if hit is None:
        if kind == "enc":
            thr, idx_min, x_limit = encode_tables(n_quantize)
            hit = (thr.to(device), idx_min, x_limit)
        else:
            hit = (decode_table(n_quantize).to(device),)
        _device_cache[k] = hit
    return hit
