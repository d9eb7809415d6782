"""Host-side logic without a GPU: the C ABI loads and exports what include/tac_b200.h declares, the
host-built constant tables (filterbank plan, mu-law levels) are right, and the module surface matches the
reference's (signatures, defaults, repr, state_dict, exceptions).  No compute call is made."""
import ctypes
import inspect
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, golden


@pytest.fixture(scope="module")
def tac():
    import torchaudio_contrib_b200 as t
    return t


def _header_functions():
    text = open(os.path.join(ROOT, "include", "tac_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tac_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(tac):
    names = _header_functions()
    assert len(names) >= 20
    lib = ctypes.CDLL(tac._cabi.LIB_PATH)
    for name in names:
        assert hasattr(lib, name), "libtac_b200.so lacks %s declared in include/tac_b200.h" % name
    assert sorted(tac._cabi.SIGNATURES) == names            # the ctypes table binds exactly the header
    assert tac._cabi.lib().tac_version() == 1


def test_num_frames_formula(tac):
    f = tac._cabi.lib().tac_stft_num_frames
    for T, n_fft, hop in [(16000, 512, 128), (160000, 2048, 512), (100000, 512, 256), (480000, 2048, 512), (100, 64, 16)]:
        assert f(T, n_fft, hop, 1) == (T + 2 * (n_fft // 2) - n_fft + hop) // hop == 1 + T // hop
        assert f(T, n_fft, hop, 0) == 1 + (T - n_fft) // hop
    assert f(10, 64, 16, 0) == 0


# ---------------------------------------------------------------------------------------- plan
def _unpack_plan(blob, n_bins, n_bands):
    """Rebuild the dense matrix from the operand images; also returns the per-slice (band_lo, n)."""
    raw = blob.numpy().tobytes()
    hdr = np.frombuffer(raw, dtype=np.int32, count=12)                      # FbPlanHeader, 48 bytes
    assert hdr[1] == n_bins and hdr[2] == n_bands
    n_chunks, n_bblocks, max_n = int(hdr[3]), int(hdr[4]), int(hdr[5])
    table = np.frombuffer(raw, dtype=np.int32, count=4 * n_chunks * n_bblocks, offset=48).reshape(n_bblocks, n_chunks, 4)
    fb = np.zeros((n_chunks * 32, n_bblocks * 128), dtype=np.float64)
    hi_only = np.zeros_like(fb)
    spans = []
    for bb in range(n_bblocks):
        for c in range(n_chunks):
            lo, n, off, _ = (int(v) for v in table[bb, c])
            spans.append((bb, c, lo, n))
            if n == 0:
                continue
            assert n % 16 == 0 and lo % 16 == 0 and lo + n <= 128 and off % 128 == 0 and n <= max_n
            img = np.frombuffer(raw, dtype=np.float32, count=2 * n * 32, offset=off).reshape(2, n // 8, 8, 8, 4)
            for j in range(n):
                for kk in range(32):
                    a, r = divmod(j, 8)
                    h = img[0, a, r, (kk >> 2) ^ r, kk & 3]
                    l = img[1, a, r, (kk >> 2) ^ r, kk & 3]
                    assert (np.float32(h).view(np.uint32) & 0x1FFF) == 0          # hi is an exact tf32
                    fb[c * 32 + kk, bb * 128 + lo + j] = float(h) + float(l)
                    hi_only[c * 32 + kk, bb * 128 + lo + j] = float(h)
    return fb[:n_bins, :n_bands], hi_only[:n_bins, :n_bands], spans


@pytest.mark.parametrize("which", ["mel16k", "dense", "wide"])
def test_filterbank_plan_images(tac, which):
    if which == "mel16k":
        fb = golden("filterbanks.npz")["fb_16k_1025x128"]
    elif which == "dense":
        fb = torch.randn(257, 36, generator=torch.Generator().manual_seed(4))
    else:
        fb = torch.randn(70, 300, generator=torch.Generator().manual_seed(5)) * (torch.rand(70, 300) > 0.9)
    lib = tac._cabi.lib()
    cap = lib.tac_fbplan_bytes(fb.size(0), fb.size(1))
    buf = torch.zeros(cap, dtype=torch.uint8)
    used = ctypes.c_int64()
    fbc = fb.contiguous()
    assert lib.tac_fbplan_build_host(fbc.data_ptr(), fb.size(0), fb.size(1), buf.data_ptr(), cap, ctypes.byref(used)) == 0
    assert 0 < used.value <= cap
    rebuilt, hi_only, spans = _unpack_plan(buf[:used.value], fb.size(0), fb.size(1))
    assert np.array_equal(rebuilt.astype(np.float32), fb.numpy())       # hi + lo is the matrix, exactly
    assert np.abs(hi_only - fb.numpy()).max() <= np.abs(fb.numpy()).max() * 2.0 ** -10
    if which == "mel16k":
        widths = [n for _, _, _, n in spans if n]
        assert max(widths) <= 48 and sum(widths) < 0.3 * 128 * len(widths)   # block skipping pays on a mel matrix
    # too-small buffer is refused, not overrun
    assert lib.tac_fbplan_build_host(fbc.data_ptr(), fb.size(0), fb.size(1), buf.data_ptr(), 64, ctypes.byref(used)) == -4
    assert b"too small" in lib.tac_last_error()


def _unpack_range_plan(blob):
    """(dense matrix rebuilt from the range plan, sum of range lengths), or None when the blob carries none."""
    raw = blob.numpy().tobytes()
    hdr = np.frombuffer(raw, dtype=np.int32, count=12)
    off, nbytes = int(hdr[8]), int(hdr[9])
    if off == 0:
        return None
    rh = np.frombuffer(raw, dtype=np.int32, count=8, offset=off)
    n_bins, n_bands, pad, nnz_pad, max_len = (int(v) for v in rh[1:6])
    assert nbytes == 32 + pad * 8 + nnz_pad * 4 and off % 16 == 0 and nbytes % 16 == 0
    meta = np.frombuffer(raw, dtype=np.int32, count=2 * pad, offset=off + 32).reshape(pad, 2)
    w = np.frombuffer(raw, dtype=np.float32, count=nnz_pad, offset=off + 32 + pad * 8)
    fb = np.zeros((n_bins, n_bands), dtype=np.float32)
    total = 0
    for b in range(pad):
        lo, ln, o = int(meta[b, 0]) & 0xffff, int(meta[b, 0]) >> 16, int(meta[b, 1])
        assert ln <= max_len and (b < n_bands or ln == 0)
        if ln:
            fb[lo:lo + ln, b] = w[o:o + ln]
            total += ln
    return fb, total


@pytest.mark.parametrize("name,fft", [("fb_16k_1025x128", 2048), ("fb_16k_513x128", 1024), ("fb_16k_129x128", 256)])
def test_range_plan_describes_the_matrix(tac, name, fft):
    """The per-band bin ranges of the fused epilogue for n_fft != 2048 (csrc/bandplan.cu build_range_plan) rebuild the
    filterbank exactly; a dense matrix gets no range plan and no fused handle."""
    g = golden("filterbanks.npz")
    if name not in g:
        fb = tac.MelFilterbank(num_freqs=fft // 2 + 1, num_mels=128, sample_rate=16000).get_filterbank()
    else:
        fb = g[name]
    lib = tac._cabi.lib()

    def build(m):
        cap = lib.tac_fbplan_bytes(m.size(0), m.size(1))
        buf = torch.zeros(cap, dtype=torch.uint8)
        used = ctypes.c_int64()
        mc = m.contiguous()
        assert lib.tac_fbplan_build_host(mc.data_ptr(), m.size(0), m.size(1), buf.data_ptr(), cap, ctypes.byref(used)) == 0
        return buf[:used.value]

    blob = build(fb)
    rebuilt, total = _unpack_range_plan(blob)
    assert np.array_equal(rebuilt, fb.numpy())
    assert total <= 2 * fb.size(0) + 2 * 128                             # a triangular bank: ~2 non-zeros per bin
    handle = lib.tac_fbplan_fused_handle(blob.data_ptr(), fft)
    assert handle != 0
    assert lib.tac_fbplan_fused_handle(blob.data_ptr(), fft * 2) == 0      # wrong fft length for this matrix
    dense = torch.randn(fft // 2 + 1, 40, generator=torch.Generator().manual_seed(9))
    dblob = build(dense)
    assert _unpack_range_plan(dblob) is None and lib.tac_fbplan_fused_handle(dblob.data_ptr(), fft) == 0


# ---------------------------------------------------------------------------------------- mu-law tables
@pytest.mark.parametrize("q", [256, 64, 2, 1024])
def test_mulaw_tables_describe_the_reference_quantiser(tac, q):
    from oracle import ref_chain as oc
    from torchaudio_contrib_b200 import _mulaw_tables as mt
    thr, idx_min, x_limit = mt.encode_tables(q)
    assert thr[0] == float("-inf") and bool((thr[1:] > thr[:-1]).all())
    g = torch.Generator().manual_seed(q)
    bits = torch.randint(-(1 << 31), (1 << 31) - 1, (300000,), generator=g, dtype=torch.int64).to(torch.int32)
    x = bits.view(torch.float32)
    x = torch.cat([x[torch.isfinite(x)], torch.rand(300000, generator=g) * 2 - 1, thr[1:], torch.nextafter(thr[1:], thr[1:] - 1)])
    want = oc.mu_law_encoding(x, q)
    got = torch.searchsorted(thr, x, right=True) - 1 + idx_min
    got = torch.where(x.abs() <= x_limit, got, torch.full_like(got, -(1 << 63)))
    assert torch.equal(got, want)
    assert torch.equal(mt.decode_table(q), oc.mu_law_decoding(torch.arange(q), q))


# ---------------------------------------------------------------------------------------- module surface
def test_signatures_match_the_reference(tac):
    def params(fn):
        return [(n, p.default) for n, p in inspect.signature(fn).parameters.items() if n != "self"]

    E = inspect.Parameter.empty
    assert params(tac.STFT.__init__) == [("fft_length", E), ("hop_length", None), ("win_length", None), ("window", None),
                                         ("center", True), ("pad_mode", "reflect"), ("normalized", False), ("onesided", True)]
    assert params(tac.stft) == [("waveforms", E)] + params(tac.STFT.__init__)
    assert params(tac.ComplexNorm.__init__) == [("power", 1.0)]
    assert params(tac.ApplyFilterbank.__init__) == [("filterbank", E)]
    assert params(tac.MelFilterbank.__init__) == [("num_freqs", 1025), ("num_mels", 128), ("min_freq", 0.0), ("max_freq", None),
                                                  ("sample_rate", None), ("htk", False)]
    assert params(tac.Spectrogram) == params(tac.STFT.__init__) + [("power", 1.0)]
    assert params(tac.Melspectrogram)[:7] == [("num_mels", 128), ("sample_rate", 22050), ("min_freq", 0.0), ("max_freq", None),
                                              ("num_freqs", None), ("htk", False), ("mel_filterbank", None)]
    assert params(tac.AmplitudeToDb.__init__) == [("ref", 1.0), ("amin", 1e-7)]
    assert params(tac.MuLawEncoding.__init__) == params(tac.MuLawDecoding.__init__) == [("n_quantize", 256)]
    assert params(tac.complex_norm) == [("complex_tensor", E), ("power", 1.0)]
    assert params(tac.create_mel_filter) == [(n, E) for n in ("num_freqs", "num_mels", "min_freq", "max_freq", "htk")]
    assert params(tac.apply_filterbank)[:2] == [("mag_specgrams", E), ("filterbank", E)]
    assert params(tac.amplitude_to_db) == [("x", E), ("ref", 1.0), ("amin", 1e-7)]
    assert params(tac.mu_law_encoding) == [("x", E), ("n_quantize", 256)]
    assert params(tac.mu_law_decoding)[:2] == [("x_mu", E), ("n_quantize", 256)]
    # rows SURVEY 8(f) N2 / N4
    assert params(tac.TimeStretch.__init__) == [("hop_length", E), ("num_freqs", E), ("fixed_rate", None)]    # layers.py:228
    assert params(tac.TimeStretch.forward) == [("complex_specgrams", E), ("overriding_rate", None)]           # layers.py:238
    assert params(tac.DbToAmplitude.__init__) == [("ref", 1.0)]                                                # layers.py:396
    assert params(tac.phase_vocoder) == [("complex_specgrams", E), ("rate", E), ("phase_advance", E)]          # functional.py:204
    assert params(tac.db_to_amplitude) == [("x", E), ("ref", 1.0)]                                             # functional.py:299
    assert params(tac.angle) == [("complex_tensor", E)] and params(tac.magphase) == [("complex_tensor", E), ("power", 1.0)]


def test_module_structure_and_state(tac):
    mel = tac.Melspectrogram(num_mels=128, sample_rate=16000, fft_length=2048, hop_length=512, num_freqs=7)
    assert isinstance(mel, torch.nn.Sequential)
    kinds = [type(m) for m in mel]                                  # iterable children (layers.py:346, test_layers.py:69)
    assert kinds == [tac.STFT, tac.ComplexNorm, tac.ApplyFilterbank]
    assert mel[1].power == 2.0 and mel[2].filterbank.shape == (1025, 128)   # num_freqs argument ignored (layers.py:330-332)
    with pytest.raises(TypeError):
        tac.Melspectrogram()                 # like the reference: Spectrogram(**kwargs) needs fft_length (layers.py:346)
    assert mel.state_dict() == {} and tac.STFT(512).state_dict() == {}
    mel.load_state_dict({}, strict=True)
    with_db = torch.nn.Sequential(*mel, tac.AmplitudeToDb())
    assert with_db.state_dict() == {}
    spec = tac.Spectrogram(512, hop_length=128)
    assert [type(m) for m in spec] == [tac.STFT, tac.ComplexNorm] and spec[1].power == 1.0
    # tests/test_layers.py:18-38
    layer = tac.STFT(fft_length=512, hop_length=256, pad_mode='reflect')
    assert torch.is_tensor(layer.window) and not layer.window.requires_grad and layer.window.size(0) <= layer.fft_length
    assert "window" in dict(layer.named_buffers())
    assert tac.STFT(512, win_length=400).window.shape == (400,)

    class Flat(tac.Filterbank):
        def __init__(self, num_freqs, num_mels, **kw):
            self.shape = (num_freqs, num_mels)

        def get_filterbank(self):
            return torch.ones(self.shape)

    custom = tac.Melspectrogram(num_mels=10, mel_filterbank=Flat, fft_length=256)
    assert custom[2].filterbank.shape == (129, 10)
    with pytest.raises(NotImplementedError):
        tac.Filterbank().get_filterbank()


def test_repr_strings(tac):
    assert repr(tac.STFT(512, 128)) == ("STFT(fft_length=512, hop_length=128, win_length=None)"
                                        "(center=True, pad_mode=reflect, normalized=False, onesided=True)")
    assert repr(tac.ComplexNorm(2.0)) == "ComplexNorm(power=2.0)"
    assert repr(tac.MelFilterbank(sample_rate=16000)) == ("MelFilterbank(num_freqs=1025, snum_mels=128, min_freq=0.0, "
                                                          "max_freq=8000), htk=False")
    assert repr(tac.AmplitudeToDb()) == "AmplitudeToDb(ref=1.0, amin=1e-07)"
    assert repr(tac.MuLawEncoding()) == "MuLawEncoding(n_quantize=256)" and repr(tac.MuLawDecoding(64)) == "MuLawDecoding(n_quantize=64)"


def test_time_stretch_and_db_to_amplitude_modules(tac):
    """tests/test_layers.py:41-52 (buffer attributes) + the reference's repr / state / error behaviour."""
    layer = tac.TimeStretch(hop_length=256, num_freqs=1025)
    assert torch.is_tensor(layer.phase_advance) and not layer.phase_advance.requires_grad
    assert layer.phase_advance.shape == (1025, 1)
    assert torch.equal(layer.phase_advance, torch.linspace(0, np.pi * 256, 1025)[..., None])       # layers.py:232-233
    assert layer.state_dict() == {} and "phase_advance" in dict(layer.named_buffers())
    assert repr(tac.TimeStretch(256, 257, fixed_rate=0.7)) == "TimeStretch(fixed_rate=0.7)"
    assert repr(tac.DbToAmplitude(ref=2.0)) == "DbToAmplitude(ref=2.0)" and tac.DbToAmplitude().state_dict() == {}
    with pytest.raises(ValueError):
        layer(torch.zeros(1, 1025, 4, 2))                                                              # layers.py:251-253
    same = torch.zeros(1, 1025, 4, 2)
    assert tac.TimeStretch(256, 1025, fixed_rate=1.0)(same) is same                                    # layers.py:257-258
    with pytest.raises(RuntimeError):
        tac.TimeStretch(256, 1025, fixed_rate=0.7)(same)                                               # CPU tensor: no fallback


def test_constructor_errors(tac):
    with pytest.raises(ValueError):
        tac.MelFilterbank()                                      # layers.py:188-190
    with pytest.raises(AssertionError):
        tac.AmplitudeToDb(ref=1e-8, amin=1e-7)                   # layers.py:366-367
    assert tac.MelFilterbank(sample_rate=22050).max_freq == 11025 and isinstance(tac.MelFilterbank(sample_rate=22050).max_freq, int)


def test_mel_matrix_is_bit_identical_to_the_reference(tac):
    g = golden("filterbanks.npz")
    assert torch.equal(tac.MelFilterbank(num_freqs=1025, num_mels=128, sample_rate=16000).get_filterbank(), g["fb_16k_1025x128"])
    assert torch.equal(tac.Melspectrogram(num_mels=128, sample_rate=48000, fft_length=2048)[2].filterbank, g["fb_48k_1025x128"])
    for fft in (256, 512, 1024, 4096):
        assert torch.equal(tac.create_mel_filter(fft // 2 + 1, 128, 0.0, 8000, False), g["fb_16k_%dx128" % (fft // 2 + 1)])
    assert torch.equal(tac.MelFilterbank(1025, 40, 30.0, None, 22050, True).get_filterbank(), g["fb_htk_22k_1025x40"])
    assert torch.equal(tac.MelFilterbank(num_freqs=257, num_mels=128, max_freq=1.0).get_filterbank(), g["fb_maxfreq1_257x128"])


def test_no_cpu_fallback(tac):
    x = torch.randn(1, 1, 4000)
    for call in (lambda: tac.Spectrogram(512, 128)(x), lambda: tac.stft(x, 512), lambda: tac.mu_law_encoding(x),
                 lambda: tac.mu_law_decoding(torch.arange(4)), lambda: tac.amplitude_to_db(x),
                 lambda: tac.complex_norm(torch.randn(4, 2)), lambda: tac.apply_filterbank(torch.randn(1, 257, 9), torch.randn(257, 5)),
                 lambda: tac.Melspectrogram(fft_length=2048)(x)):
        with pytest.raises(RuntimeError, match="CUDA"):
            call()
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="CUDA"):
            tac.HostPipeline(512, 128)


def test_missing_library_fails_loudly(tac, tmp_path):
    with pytest.raises(ImportError, match="no CPU fallback"):
        saved = tac._cabi._lib
        tac._cabi._lib = None
        try:
            tac._cabi.load(str(tmp_path / "absent.so"))
        finally:
            tac._cabi._lib = saved


# ---------------------------------------------------------------------------------------- backward / multi-GPU host logic
def test_backward_workspace_sizes_and_argument_checks(tac):
    """Workspace formulas of the adjoint entry points, and that a too-small workspace is refused before any launch."""
    lib = tac._cabi.lib()
    frames = lib.tac_stft_num_frames(160000, 2048, 512, 1)
    assert frames == 313
    rows = 64 * frames
    assert lib.tac_stft_backward_workspace_bytes(64, 160000, 2048, 512, 1) == rows * 2048 * 4 + rows * 1056 * 4
    assert lib.tac_stft_backward_workspace_bytes(64, 160000, 512, 128, 1) == 64 * lib.tac_stft_num_frames(160000, 512, 128, 1) * 512 * 4
    assert lib.tac_stft_backward_workspace_bytes(0, 160000, 512, 128, 1) == 0
    a = (rows * 1056 * 4 + 255) // 256 * 256
    assert lib.tac_melspec_backward_workspace_bytes(64, 160000, 2048, 512, 1) == a + rows * 2048 * 4 + 32768
    f1k = lib.tac_stft_num_frames(160000, 1024, 256, 1)
    a = (64 * f1k * 513 * 4 + 255) // 256 * 256
    assert lib.tac_melspec_backward_workspace_bytes(64, 160000, 1024, 256, 1) == a + 64 * f1k * 1024 * 4 + 32768
    # argument validation happens on the host, before any CUDA call: fake non-null pointers are never dereferenced
    fake = ctypes.c_void_p(4096)
    rc = lib.tac_melspec_backward_f32(fake, 2, 8000, 8000, fake, 2048, 512, 1, 0, 0, 2.0, fake, 128, fake, 128 * 16, 16, 1,
                                      fake, fake, 1024, None)
    assert rc == tac._cabi.TAC_ERR_WORKSPACE
    assert b"tac_melspec_backward_workspace_bytes" in lib.tac_last_error()
    rc = lib.tac_melspec_backward_f32(fake, 2, 900, 900, fake, 2048, 512, 1, 0, 0, 2.0, fake, 128, fake, 128, 1, 1, fake, fake, 1 << 30, None)
    assert rc == tac._cabi.TAC_ERR_INVALID and b"Padding size" in lib.tac_last_error()       # reflect pad >= length, like the forward


def test_one_kernel_path_predicate(tac, monkeypatch):
    """`_mel_frame_major`: which calls write the frame-major buffer (one-kernel path, reference layout)."""
    F = tac.functional
    tri = tac.MelFilterbank(num_freqs=1025, num_mels=128, sample_rate=16000).get_filterbank()
    dense = torch.rand(1025, 24) + 0.1
    cpu = torch.device("cpu")
    assert F._mel_frame_major(tri, 2048, "reference", cpu, None) is True
    assert F._mel_frame_major(tri, 2048, "contiguous", cpu, None) is False
    assert F._mel_frame_major(dense, 2048, "reference", cpu, None) is False                 # tensor-core path
    tri_small = tac.MelFilterbank(num_freqs=513, num_mels=64, sample_rate=16000).get_filterbank()
    assert F._mel_frame_major(tri_small, 1024, "reference", cpu, None) is True               # range-plan epilogue (round 2)
    assert F._mel_frame_major(torch.rand(513, 24) + 0.1, 1024, "reference", cpu, None) is False
    tri_4096 = tac.MelFilterbank(num_freqs=2049, num_mels=64, sample_rate=16000).get_filterbank()
    assert F._mel_frame_major(tri_4096, 4096, "reference", cpu, None) is True
    tri_8192 = tac.MelFilterbank(num_freqs=4097, num_mels=64, sample_rate=16000).get_filterbank()
    assert F._mel_frame_major(tri_8192, 8192, "reference", cpu, None) is False               # two-kernel path
    monkeypatch.setenv("TAC_MELSPEC_FUSED", "0")
    assert F._mel_frame_major(tri, 2048, "reference", cpu, None) is False


def test_gradient_dispatch_helpers(tac):
    F = tac.functional
    x = torch.zeros(4, requires_grad=True)
    assert F._wants_grad(x) and not F._wants_grad(x.detach()) and not F._wants_grad(None)
    with torch.no_grad():
        assert not F._wants_grad(x)
    F._no_param_grad(torch.zeros(3), "window")                                             # constants without grad are fine
    with pytest.raises(RuntimeError, match="is a constant on this path"):
        F._no_param_grad(x, "window")


def test_shipped_mulaw_tables_equal_fresh_bisection():
    """tac_mulaw_tables_host(256): the decision levels / decoded values compiled into the library (csrc/mulaw_table256.inc,
    what a non-Python caller of the C ABI gets) are the ones a fresh bisection with the reference's fp32 torch CPU chain
    finds on this host; other n_quantize come from the host libm and say so (*exact = 0)."""
    import ctypes
    from torchaudio_contrib_b200 import _cabi, _mulaw_tables
    lib = _cabi.lib()

    def tables(nq):
        n, imin, xl, ex = ctypes.c_int(), ctypes.c_int(), ctypes.c_float(), ctypes.c_int()
        _cabi.check(lib.tac_mulaw_tables_host(nq, None, 0, ctypes.byref(n), ctypes.byref(imin), ctypes.byref(xl), None, ctypes.byref(ex)))
        thr, dec = torch.empty(n.value), torch.empty(nq)
        _cabi.check(lib.tac_mulaw_tables_host(nq, thr.data_ptr(), n.value, ctypes.byref(n), ctypes.byref(imin), ctypes.byref(xl),
                                              dec.data_ptr(), ctypes.byref(ex)))
        return thr, imin.value, xl.value, dec, ex.value

    thr, imin, xl, dec, exact = tables(256)
    want_thr, want_imin, want_xl = _mulaw_tables.encode_tables(256)
    assert exact == 1 and imin == want_imin and xl == want_xl
    assert torch.equal(thr, want_thr) and torch.equal(dec, _mulaw_tables.decode_table(256))
    thr, imin, xl, dec, exact = tables(64)
    want_thr, want_imin, want_xl = _mulaw_tables.encode_tables(64)
    assert exact == 0 and imin == want_imin and xl == want_xl and thr.numel() == want_thr.numel()
    close = (thr[1:].view(torch.int32) - want_thr[1:].view(torch.int32)).abs().max().item()
    assert close <= 64                                         # libm vs torch log1p: the levels agree to a few ulp
    n = ctypes.c_int()
    with pytest.raises(_cabi.TacError):
        _cabi.check(lib.tac_mulaw_tables_host(256, thr.data_ptr(), 10, ctypes.byref(n), ctypes.byref(n), ctypes.byref(ctypes.c_float()), None, None))


def test_filterbank_plan_cache_identity(tac):
    """The plan cache (functional._plan_for): the same matrix and its aliases hit one plan, an in-place edit or another
    tensor does not; a cached plan keeps the matrix's storage alive so its address cannot be recycled under the key."""
    F = tac.functional
    F.invalidate_filterbank_plans()
    fb = tac.MelFilterbank(num_freqs=257, num_mels=40, sample_rate=16000).get_filterbank()
    p1 = F._plan_for(fb, "cpu")
    assert F._plan_for(fb, "cpu") is p1
    assert F._plan_for(fb.detach(), "cpu") is p1                # what the backward pass hands in
    fb.mul_(2.0)                                                # bumps the version counter
    p2 = F._plan_for(fb, "cpu")
    assert p2 is not p1
    other = fb.clone()
    assert F._plan_for(other, "cpu") is not p2
    cache = {}
    q1 = F._plan_for(fb, "cpu", cache)
    assert F._plan_for(fb, "cpu", cache) is q1 and F._plan_for(other, "cpu", cache) is not q1
    for _ in range(20):                                         # bounded: oldest plans are evicted
        F._plan_for(torch.rand(33, 8), "cpu")
    assert len(F._GLOBAL_PLANS) <= F._GLOBAL_PLANS_MAX
    F.invalidate_filterbank_plans()
    assert len(F._GLOBAL_PLANS) == 0


def test_c_example_compiles_against_the_header():
    """examples/c_abi_mulaw.c is plain C against include/tac_b200.h: it must compile and link here (no GPU needed to build)."""
    import shutil
    import subprocess
    import tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None or not os.path.exists("/usr/local/cuda/include/cuda_runtime.h"):
        pytest.skip("no C compiler / CUDA headers here")
    libdir = os.path.join(root, "torchaudio_contrib_b200", "lib")
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run([cc, "-O2", "-Wall", "-I", os.path.join(root, "include"), "-I", "/usr/local/cuda/include",
                        os.path.join(root, "examples", "c_abi_mulaw.c"), "-o", os.path.join(tmp, "c_abi_mulaw"), "-L", libdir,
                        "-ltac_b200", "-L", "/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath," + libdir], check=True)
