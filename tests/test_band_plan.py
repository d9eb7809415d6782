"""Band plan (csrc/bandplan.cu) checked on the CPU: the host builder runs here through the C ABI and a
numpy model of the fused kernel's band_contract step (csrc/stft.cu) consumes its output.

The model follows the kernel statement for statement -- stash layout (row stride 33), lane l <- bins
32 l .. 32 l + 31, the mask-driven `row[i] = u; u = v; v = 0` steps, the end-of-run pair at `end_pos`, the
per-band lists -- so a layout mistake in the plan or in the kernel's index arithmetic shows up here, without a
GPU.  The GPU parity tests (test_gpu_parity.py) then pin the kernel itself against the oracle.
"""
import ctypes

import numpy as np
import pytest
import torch

from torchaudio_contrib_b200 import _cabi
from torchaudio_contrib_b200 import functional as F

STRIDE, NYQ, ZERO, FLOATS, SHIFT_ROW = 33, 32 * 33 - 1, 1060, 1064, 17
OFF_W, OFF_META, OFF_COMB = 32, 32 + 8192, 32 + 8192 + 512


def build_plan(fb):
    lib = _cabi.lib()
    fb = np.ascontiguousarray(fb, dtype=np.float32)
    n_bins, n_bands = fb.shape
    cap = int(lib.tac_fbplan_bytes(n_bins, n_bands))
    host = np.zeros(cap, dtype=np.uint8)
    used = ctypes.c_int64(0)
    _cabi.check(lib.tac_fbplan_build_host(fb.ctypes.data, n_bins, n_bands, host.ctypes.data, cap, ctypes.byref(used)))
    handle = int(lib.tac_fbplan_band_handle(host.ctypes.data))
    return host[:used.value], handle


def kernel_model(blob, handle, power_rows, n_bands):
    """band_contract of csrc/stft.cu in numpy float32, one frame at a time."""
    off, cmax = handle & ((1 << 48) - 1), handle >> 48
    bp = blob[off:]
    hdr = bp[:32].view(np.int32)
    assert hdr[1] == 1025 and hdr[2] == n_bands and hdr[4] == cmax and hdr[6] == ZERO
    pad = int(hdr[5])
    w = bp[OFF_W:OFF_W + 8192].view(np.float32).reshape(16, 32, 4)
    meta = bp[OFF_META:OFF_META + 512].view(np.uint32).reshape(32, 4)
    comb = bp[OFF_COMB:OFF_COMB + cmax * pad * 2].view(np.uint16).reshape(cmax, pad)
    fast_off = int(hdr[7])
    assert (fast_off != 0) == (n_bands <= 128 and cmax <= 4)
    if fast_off:
        assert fast_off == (OFF_COMB + cmax * pad * 2 + 15) // 16 * 16          # where pipeline.cu expects it
        fast = bp[fast_off:fast_off + 2048].view(np.uint32).reshape(32, 4, 4)
        # the same lists once more for the frame-major layout: lane l holds the adjacent bands 4 l .. 4 l + 3
        quad = bp[fast_off + 2048:fast_off + 4096].view(np.uint32).reshape(32, 4, 4)
        for m in range(128):
            assert (quad[m // 4, m % 4] == fast[m % 32, m // 32]).all()
    out = np.zeros((power_rows.shape[0], n_bands), dtype=np.float32)
    for f, p in enumerate(power_rows.astype(np.float32)):
        stash = np.full(FLOATS, np.nan, dtype=np.float32)
        stash[ZERO] = 0.0
        for k1 in range(32):                                   # emit: lane = column, k1 = row
            stash[k1 * STRIDE + 1:k1 * STRIDE + 32] = p[32 * k1 + 1:32 * k1 + 32]
            stash[k1 * STRIDE - (1 if k1 >= SHIFT_ROW else 0)] = p[32 * k1]     # column 0 of rows >= 17 sits one float lower
        stash[NYQ] = p[1024]
        pw = np.stack([stash[l * STRIDE:l * STRIDE + 32].copy() for l in range(32)])
        for l in range(SHIFT_ROW, 32):
            pw[l, 0] = stash[l * STRIDE - 1]
        p_last = stash[NYQ]
        stash[:] = np.nan                                      # anything read later must have been stored
        stash[ZERO] = 0.0
        for l in range(32):
            mask, end_pos = int(meta[l, 0]), int(meta[l, 1])
            e0, e1 = meta[l, 2:4].view(np.float32)
            u = v = np.float32(0)
            for i in range(32):
                wi = w[i // 2, l, 2 * (i % 2):2 * (i % 2) + 2]
                u = np.float32(u + pw[l, i] * wi[0])
                v = np.float32(v + pw[l, i] * wi[1])
                if mask >> i & 1:
                    stash[l * STRIDE + i] = u
                    u, v = v, np.float32(0)
            last = p_last if l == 31 else np.float32(0)
            u = np.float32(u + last * e0)
            v = np.float32(v + last * e1)
            stash[end_pos], stash[end_pos + 1] = u, v
        for m in range(n_bands):
            acc = np.float32(0)
            for c in range(cmax):
                acc = np.float32(acc + stash[comb[c, m]])
            out[f, m] = acc
            if fast_off:                                       # the per-lane 4 x 4 form must give the same sum
                e = fast[m % 32, m // 32]
                assert (e % 4 == 0).all()
                a, b, c, d = (stash[i // 4] for i in e)
                assert np.float32(np.float32(np.float32(a + b) + c) + d) == acc
        if fast_off:
            for m in range(n_bands, 128):
                assert (fast[m % 32, m // 32] == ZERO * 4).all()
    return out


MEL_CASES = [(128, 16000, False), (128, 48000, False), (128, 22050, True), (40, 16000, False), (64, 8000, True),
             (80, 44100, False)]


@pytest.mark.parametrize("num_mels,sr,htk", MEL_CASES)
def test_mel_matrices_have_a_band_plan_and_the_kernel_model_matches_matmul(num_mels, sr, htk):
    fb = F.create_mel_filter(1025, num_mels, 0.0, sr // 2, htk).numpy()
    blob, handle = build_plan(fb)
    assert handle != 0, "a triangular mel matrix must take the fused path"
    rng = np.random.default_rng(num_mels + sr)
    power = (rng.standard_normal((3, 1025)) ** 2 * 1e3).astype(np.float32)
    power[1, 100:140] *= 1e6                                   # a strong partial next to weak bins
    got = kernel_model(blob, handle, power, num_mels)
    want = power.astype(np.float64) @ fb.astype(np.float64)
    assert not np.isnan(got).any()
    np.testing.assert_allclose(got, want, rtol=2e-6, atol=0)


def test_min_freq_and_narrow_band_ranges():
    fb = F.create_mel_filter(1025, 96, 300.0, 6000.0, False).numpy()
    blob, handle = build_plan(fb)
    assert handle != 0
    power = np.abs(np.random.default_rng(5).standard_normal((2, 1025))).astype(np.float32)
    np.testing.assert_allclose(kernel_model(blob, handle, power, 96), power.astype(np.float64) @ fb, rtol=2e-6)


def test_last_bin_weight_reaches_lane_31():
    # a chain whose last band peaks at bin 1024: exercises the extra bin of lane 31 and its mask bit 31
    n_bands = 8
    fb = np.zeros((1025, n_bands), dtype=np.float32)
    edges = np.linspace(0, 1024, n_bands + 1).round().astype(int)
    edges[-2] = 1023                                           # band 6 -> 7 steps exactly between bins 1023 and 1024
    for b in range(n_bands):
        lo, hi = edges[b], edges[b + 1]
        for k in range(lo, hi + (1 if b == n_bands - 1 else 0)):
            t = (k - lo) / max(hi - lo, 1)
            fb[k, b] = 1.0 - 0.5 * t
            if b + 1 < n_bands:
                fb[k, b + 1] = 0.5 * t if t > 0 else 0.0
    assert fb[1024, n_bands - 1] != 0
    blob, handle = build_plan(fb)
    assert handle != 0
    power = np.abs(np.random.default_rng(9).standard_normal((2, 1025))).astype(np.float32) + 1
    np.testing.assert_allclose(kernel_model(blob, handle, power, n_bands), power.astype(np.float64) @ fb, rtol=2e-6)


def test_dense_and_gapped_matrices_fall_back_to_the_tensor_core_plan():
    rng = np.random.default_rng(0)
    _, handle = build_plan(rng.standard_normal((1025, 120)).astype(np.float32))
    assert handle == 0
    _, handle = build_plan(rng.standard_normal((257, 36)).astype(np.float32))
    assert handle == 0
    fb = F.create_mel_filter(1025, 128, 0.0, 8000, False).numpy().copy()
    fb[500, 3] = 0.25                                          # a third non-zero far from the diagonal
    _, handle = build_plan(fb)
    assert handle == 0
    # other fft sizes keep the two-kernel pipeline
    _, handle = build_plan(F.create_mel_filter(513, 64, 0.0, 8000, False).numpy())
    assert handle == 0


def test_skipped_bands_are_rejected_or_exact():
    # 2048-point bins but so many mel bands that some triangles contain no bin at all (the 13 empty bands of
    # SURVEY H4 at fft 256): the chain b -> b + 1 breaks; whichever way the builder decides, a plan it returns
    # must reproduce the matmul
    fb = F.create_mel_filter(1025, 700, 0.0, 8000, False).numpy()
    blob, handle = build_plan(fb)
    if handle:
        power = np.abs(np.random.default_rng(1).standard_normal((1, 1025))).astype(np.float32)
        np.testing.assert_allclose(kernel_model(blob, handle, power, 700), power.astype(np.float64) @ fb, rtol=2e-6,
                                   atol=1e-30)


# ------------------------------------------------------------------------------- index logic of stft2048_kernel, modelled
def test_contiguous_chunks_partition_the_frames():
    """csrc/stft.cu: CTA b takes frames [chunk0, chunk1) with chunk sizes differing by at most one."""
    for n_all, grid in [(20032, 148), (480256, 148), (7, 148), (148, 148), (149, 148), (2368, 148), (1, 1), (33, 3)]:
        per, extra = divmod(n_all, grid)
        covered = []
        for b in range(grid):
            c0 = b * per + min(b, extra)
            c1 = c0 + per + (1 if b < extra else 0)
            covered.extend(range(c0, c1))
            assert c1 - c0 in (per, per + 1)
        assert covered == list(range(n_all))


def test_stash_stores_are_bank_conflict_free():
    """The two store instructions of one untangling step (csrc/stft.cu emit2): `dst` = (row k1, column lane) and the
    mirror = (row 31 - k1, column 32 - lane), lane 0's mirror at (row 32 - k1, column 0) moved one float down
    (bandplan.cuh kStashShiftRow).  Every lane must hit its own bank, and every bin its own float."""
    owner = {}
    for k1 in range(16):
        dst = [k1 * STRIDE + lane for lane in range(32)]
        mir = [(STRIDE - 1 if lane == 0 else 32 - lane) + (31 - k1) * STRIDE for lane in range(32)]
        for addrs in (dst, mir):
            assert len({a % 32 for a in addrs}) == 32
        for lane in range(32):
            k = 32 * k1 + lane
            owner[dst[lane]] = k
            owner[mir[lane]] = 1024 - k
    owner[16 * STRIDE] = 512                                  # lane 0, step k1 = 16 (its own mirror)
    assert len(owner) == 1025 and sorted(owner.values()) == list(range(1025))
    # the consumer's view: lane l reads bins 32 l + i at l * 33 + i, column 0 of rows >= 17 one float lower
    for lane in range(32):
        for i in range(32):
            addr = lane * STRIDE + i - (1 if (i == 0 and lane >= SHIFT_ROW) else 0)
            assert owner[addr] == 32 * lane + i
    assert owner[NYQ] == 1024
