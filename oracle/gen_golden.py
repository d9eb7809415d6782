"""Generate tests/golden/*.npz from the UNMODIFIED reference.  Build-container only.

    python oracle/gen_golden.py            # needs /root/reference (read-only checkout)

The reference is imported by path; the only harness-side adaptation is a `torch.stft` shim
that requests `return_complex=True` and returns `view_as_real` (the legacy layout the
reference was written against -- torch 2.x refuses the legacy call).  No reference file is
touched or copied.  For every fixture the script also asserts that `oracle.ref_chain`
reproduces the reference output bit for bit (`torch.equal`), which is what pins the oracle.
"""
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("TAC_REFERENCE_PATH", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")


def _load_reference():
    sys.path.insert(0, REF)
    native_stft = torch.stft

    def legacy_stft(*args, **kwargs):
        kwargs["return_complex"] = True
        return torch.view_as_real(native_stft(*args, **kwargs))

    torch.stft = legacy_stft
    import torchaudio_contrib as ref      # noqa: E402  (the reference package)
    torch.stft = native_stft              # module-level name is re-resolved at call time,
    ref._legacy_stft = legacy_stft        # so keep a switch for the calls below
    ref._native_stft = native_stft
    return ref


class _shimmed:
    """Context manager: torch.stft -> legacy layout while the reference runs."""

    def __init__(self, ref):
        self.ref = ref

    def __enter__(self):
        torch.stft = self.ref._legacy_stft

    def __exit__(self, *exc):
        torch.stft = self.ref._native_stft


def _same(a, b, what):
    if not torch.equal(a, b):
        raise SystemExit("oracle.ref_chain differs from the reference on %s (max abs %g)"
                         % (what, (a.double() - b.double()).abs().max().item()))


def _save(name, **arrays):
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name), **{k: np.ascontiguousarray(v) for k, v in arrays.items()})
    print("wrote", name, {k: tuple(np.shape(v)) for k, v in arrays.items()})


def widen(ref, oc):
    """Fixtures of the rows SURVEY 8(f) marks "next": N2 phase_vocoder / TimeStretch, N4 db_to_amplitude,
    angle, magphase.  Own seed, own files: `python oracle/gen_golden.py widen` leaves the others untouched."""
    g = torch.Generator().manual_seed(20261017)

    def randn(*shape, dtype=torch.float32):
        return torch.randn(*shape, generator=g, dtype=dtype)

    # ---- angle / magphase / db_to_amplitude -------------------------------------------------
    z = randn(3, 129, 40, 2)
    z[0, 0, 0] = torch.tensor([0.0, 0.0])
    z[0, 0, 1] = torch.tensor([-1.0, 0.0])
    z[0, 0, 2] = torch.tensor([-1.0, -0.0])
    ang = ref.angle(z)
    _same(ang, oc.angle(z), "angle")
    mag, ph = ref.magphase(z, 2.0)
    m2, p2 = oc.magphase(z, 2.0)
    _same(mag, m2, "magphase mag")
    _same(ph, p2, "magphase phase")
    db = torch.cat([torch.tensor([-60.0, -40.0, -10.0, 0.0, 10.0, 60.0]), randn(500) * 30])
    amp = ref.db_to_amplitude(db, ref=1.0)
    _same(amp, oc.db_to_amplitude(db, 1.0), "db_to_amplitude")
    amp3 = ref.db_to_amplitude(db, ref=3.0)
    _same(amp3, oc.db_to_amplitude(db, 3.0), "db_to_amplitude ref 3")
    _save("pointwise_next.npz", z=z.numpy(), angle=ang.numpy(), mag_p2=mag.numpy(), db=db.numpy(), amp_ref1=amp.numpy(),
          amp_ref3=amp3.numpy())

    # ---- phase vocoder: the reference in float64 (its own test's precision) and in float32 -----
    hop, bins, frames = 256, 65, 90
    spec32 = randn(2, 2, bins, frames, 2)
    blob = {"spec": spec32.numpy(), "hop": np.int64(hop)}
    for rate in (0.5, 1.01, 1.3, 2.0):
        prior = torch.get_default_dtype()
        torch.set_default_dtype(torch.float64)                 # as tests/test_functional.py:76-93 does
        try:
            spec64 = spec32.double()
            adv64 = torch.linspace(0, np.pi * hop, bins)[..., None]
            y64 = ref.phase_vocoder(spec64, rate, adv64)
            _same(y64, oc.phase_vocoder(spec64, rate, adv64), "phase_vocoder f64 rate %g" % rate)
        finally:
            torch.set_default_dtype(prior)
        adv32 = torch.linspace(0, np.pi * hop, bins)[..., None]
        y32 = ref.phase_vocoder(spec32, rate, adv32)
        _same(y32, oc.phase_vocoder(spec32, rate, adv32), "phase_vocoder f32 rate %g" % rate)
        tag = ("%g" % rate).replace(".", "p")
        blob["out64_" + tag] = y64.numpy()
        blob["out32_" + tag] = y32.numpy()
    _save("phase_vocoder.npz", **blob)
    print("widening fixtures reproduce under oracle.ref_chain")


def grads(ref, oc):
    """Fixtures of the backward passes: d loss / d input of the UNMODIFIED reference under torch
    autograd for a fixed upstream gradient, and the check that oracle.ref_chain's autograd gives the same bits.
    `python oracle/gen_golden.py grads` writes tests/golden/grads.npz only."""
    g = torch.Generator().manual_seed(20261018)

    def randn(*shape):
        return torch.randn(*shape, generator=g)

    blob = {}

    def record(tag, x, run_ref, run_oc):
        xr = x.clone().requires_grad_(True)
        with _shimmed(ref):
            yr = run_ref(xr)
        gy = randn(*yr.shape)
        (gxr,) = torch.autograd.grad(yr, xr, gy)
        xo = x.clone().requires_grad_(True)
        yo = run_oc(xo)
        (gxo,) = torch.autograd.grad(yo, xo, gy)
        _same(yr.detach(), yo.detach(), "grads/%s forward" % tag)
        _same(gxr, gxo, "grads/%s backward" % tag)
        blob[tag + "_x"] = x.numpy()
        blob[tag + "_gy"] = gy.numpy()
        blob[tag + "_gx"] = gxr.numpy()

    x = randn(2, 2, 3000)
    stft_cases = {
        "stft_512_128": dict(fft_length=512, hop_length=128),
        "stft_winlen_norm": dict(fft_length=256, hop_length=64, win_length=200, normalized=True),
        "stft_nocenter": dict(fft_length=512, hop_length=100, center=False),
        "stft_constant": dict(fft_length=256, hop_length=64, pad_mode='constant'),
        "stft_replicate": dict(fft_length=256, hop_length=64, pad_mode='replicate'),
        "stft_circular": dict(fft_length=256, hop_length=64, pad_mode='circular'),
        "stft_twosided": dict(fft_length=128, hop_length=32, onesided=False),
        "stft_2048": dict(fft_length=2048, hop_length=512),
    }
    for tag, kw in stft_cases.items():
        win = torch.hann_window(kw.get("win_length", kw["fft_length"]))
        record(tag, x, lambda t, kw=kw, win=win: ref.stft(t, window=win, **kw), lambda t, kw=kw, win=win: oc.stft(t, window=win, **kw))
    for power in (1.0, 2.0, 0.7):
        tag = "spec_p%s" % ("%g" % power).replace(".", "_")
        record(tag, x, lambda t, power=power: ref.Spectrogram(fft_length=512, hop_length=128, power=power)(t),
               lambda t, power=power: oc.spectrogram(t, 512, 128, power=power))
    record("spec_twosided", x, lambda t: ref.Spectrogram(fft_length=128, hop_length=32, onesided=False, power=2.0)(t),
           lambda t: oc.spectrogram(t, 128, 32, onesided=False, power=2.0))

    x = randn(2, 2, 12000)
    record("mel_2048", x, lambda t: ref.Melspectrogram(num_mels=128, sample_rate=16000, fft_length=2048, hop_length=512)(t),
           lambda t: oc.melspectrogram(t, 128, 16000, fft_length=2048, hop_length=512))
    db = ref.AmplitudeToDb()
    record("mel_2048_db", x,
           lambda t: db(ref.Melspectrogram(num_mels=128, sample_rate=48000, fft_length=2048, hop_length=512)(t)),
           lambda t: oc.melspectrogram(t, 128, 48000, to_db=True, fft_length=2048, hop_length=512))
    record("mel_1024_64", x, lambda t: ref.Melspectrogram(num_mels=64, sample_rate=22050, fft_length=1024, hop_length=256)(t),
           lambda t: oc.melspectrogram(t, 64, 22050, fft_length=1024, hop_length=256))

    # the separate stages
    z = randn(2, 65, 40, 2)
    z[0, 0, 0] = 0.0                                                        # |z| = 0: torch.norm's subgradient is 0
    for power in (1.0, 2.0, 0.5):
        tag = "cnorm_p%s" % ("%g" % power).replace(".", "_")
        if power < 1.0:
            zz = z.clone()
            zz[0, 0, 0] = 1.0                                               # 0^(p-1) is inf for p < 1: keep the fixture finite
        else:
            zz = z
        record(tag, zz, lambda t, power=power: ref.complex_norm(t, power), lambda t, power=power: oc.complex_norm(t, power))
    fb = randn(65, 20)
    blob["fb_dense"] = fb.numpy()
    spec = randn(3, 2, 65, 50).abs()
    record("fbank_dense", spec, lambda t: ref.apply_filterbank(t, fb), lambda t: oc.apply_filterbank(t, fb))
    amp = randn(4, 30, 70)
    amp[0, 0, :8] = torch.tensor([0.0, 1e-5, -1e-5, 3.1622776e-4, 3.2e-4, -3.2e-4, 1.0, -2.0])   # around sqrt(amin)
    record("todb", amp, lambda t: ref.amplitude_to_db(t, ref=1.0, amin=1e-7), lambda t: oc.amplitude_to_db(t, 1.0, 1e-7))
    _save("grads.npz", **blob)
    print("gradient fixtures reproduce under oracle.ref_chain autograd")


def grads_more(ref, oc):
    """Round-2 gradient fixtures: the pointwise operators of SURVEY 8f N4 (db_to_amplitude, angle, magphase, float
    mu_law_decoding) and the gradient w.r.t. a learnable filterbank (apply_filterbank, Melspectrogram chain), from the
    UNMODIFIED reference under torch autograd.  Own seed, own file: `python oracle/gen_golden.py grads_more` writes
    tests/golden/grads_more.npz only."""
    g = torch.Generator().manual_seed(20261017)

    def randn(*shape):
        return torch.randn(*shape, generator=g)

    blob = {}

    def record(tag, inputs, run_ref, run_oc):
        """inputs: dict name -> tensor (all differentiated); outputs may be a tensor or a tuple of tensors."""
        def run(fn):
            leaves = {k: v.clone().requires_grad_(True) for k, v in inputs.items()}
            out = fn(**leaves)
            return leaves, (out if isinstance(out, tuple) else (out,))
        with _shimmed(ref):
            lr, yr = run(run_ref)
        gys = [randn(*y.shape) for y in yr]
        gr = torch.autograd.grad(yr, list(lr.values()), gys)
        lo, yo = run(run_oc)
        go = torch.autograd.grad(yo, list(lo.values()), gys)
        for a, b in zip(yr, yo):
            _same(a.detach(), b.detach(), "grads_more/%s forward" % tag)
        for a, b in zip(gr, go):
            _same(a, b, "grads_more/%s backward" % tag)
        for k, v in inputs.items():
            blob["%s_in_%s" % (tag, k)] = v.numpy()
        for i, gy in enumerate(gys):
            blob["%s_gy%d" % (tag, i)] = gy.numpy()
        for k, gk in zip(inputs, gr):
            blob["%s_g_%s" % (tag, k)] = gk.numpy()

    db = randn(3, 40, 50) * 20.0
    record("fromdb", dict(x=db), lambda x: ref.db_to_amplitude(x, ref=2.0), lambda x: oc.db_to_amplitude(x, 2.0))
    z = randn(2, 33, 40, 2)
    record("angle", dict(z=z), lambda z: ref.angle(z), lambda z: oc.angle(z))
    record("magphase_p1", dict(z=z), lambda z: ref.magphase(z, 1.0), lambda z: oc.magphase(z, 1.0))
    record("magphase_p2", dict(z=z), lambda z: ref.magphase(z, 2.0), lambda z: oc.magphase(z, 2.0))
    codes = torch.rand(4, 1000, generator=g) * 255.0
    record("mudec", dict(c=codes), lambda c: ref.mu_law_decoding(c, 256), lambda c: oc.mu_law_decoding(c, 256))
    spec = randn(3, 2, 65, 50).abs()
    fb = randn(65, 20)
    record("fbank_param", dict(s=spec, fb=fb), lambda s, fb: ref.apply_filterbank(s, fb), lambda s, fb: oc.apply_filterbank(s, fb))
    x = randn(2, 1, 9000)
    mel_fb = oc.mel_filterbank_for(64, 16000, fft_length=2048)
    amp = ref.AmplitudeToDb()

    def chain_ref(x, fb):
        return amp(ref.apply_filterbank(ref.Spectrogram(fft_length=2048, hop_length=512, power=2.0)(x), fb))

    def chain_oc(x, fb):
        return oc.amplitude_to_db(oc.apply_filterbank(oc.spectrogram(x, 2048, 512, power=2.0), fb), 1.0, 1e-7)

    record("melchain_param", dict(x=x, fb=mel_fb), chain_ref, chain_oc)
    _save("grads_more.npz", **blob)
    print("round-2 gradient fixtures reproduce under oracle.ref_chain autograd")


def grads_pv(ref, oc):
    """Phase-vocoder gradient fixtures (functional.py:204-274 under torch autograd), float64 with float64 as torch's default
    dtype like the reference's own value test (tests/test_functional.py:76-93): stretch and compress rates, a rate with
    several output steps per input frame.  `python oracle/gen_golden.py grads_pv` writes tests/golden/grads_pv.npz only."""
    g = torch.Generator().manual_seed(20261018)
    blob = {}
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        for tag, rate, shape, hop in (("r07", 0.7, (2, 33, 41, 2), 16), ("r13", 1.3, (1, 2, 65, 50, 2), 32), ("r20", 2.0, (3, 17, 37, 2), 8),
                                      ("r03", 0.3, (1, 9, 12, 2), 4)):
            z = torch.randn(*shape, generator=g, dtype=torch.float64)
            adv = torch.linspace(0, math.pi * hop, shape[-3], dtype=torch.float64)[..., None]
            zr = z.clone().requires_grad_(True)
            with _shimmed(ref):
                yr = ref.phase_vocoder(zr, rate, adv)
            gy = torch.randn(*yr.shape, generator=g, dtype=torch.float64)
            (gr,) = torch.autograd.grad(yr, zr, gy)
            zo = z.clone().requires_grad_(True)
            yo = oc.phase_vocoder(zo, rate, adv)
            (go,) = torch.autograd.grad(yo, zo, gy)
            _same(yr.detach(), yo.detach(), "grads_pv/%s forward" % tag)
            _same(gr, go, "grads_pv/%s backward" % tag)
            blob[tag + "_z"], blob[tag + "_adv"], blob[tag + "_gy"], blob[tag + "_gz"], blob[tag + "_y"] = \
                z.numpy(), adv.numpy(), gy.numpy(), gr.numpy(), yr.detach().numpy()
            blob[tag + "_rate"] = np.array(rate)
    finally:
        torch.set_default_dtype(old)
    _save("grads_pv.npz", **blob)
    print("phase-vocoder gradient fixtures reproduce under oracle.ref_chain autograd")


def grads_win(ref, oc):
    """Gradients w.r.t. a learnable window (functional.py:93-107 under torch autograd) from the UNMODIFIED reference: the
    complex stft (full-length and win_length < fft_length windows, normalized, no centring), the Spectrogram chain and the
    mel + dB chain, each with the waveform's gradient alongside.  `python oracle/gen_golden.py grads_win` writes
    tests/golden/grads_win.npz only."""
    g = torch.Generator().manual_seed(20261019)

    def randn(*shape):
        return torch.randn(*shape, generator=g)

    blob = {}

    def record(tag, x, w, run_ref, run_oc):
        def run(fn):
            xl, wl = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
            return (xl, wl), fn(xl, wl)
        with _shimmed(ref):
            lr, yr = run(run_ref)
        gy = randn(*yr.shape)
        gr = torch.autograd.grad(yr, lr, gy)
        lo, yo = run(run_oc)
        go = torch.autograd.grad(yo, lo, gy)
        _same(yr.detach(), yo.detach(), "grads_win/%s forward" % tag)
        for a, b in zip(gr, go):
            _same(a, b, "grads_win/%s backward" % tag)
        blob[tag + "_x"], blob[tag + "_w"], blob[tag + "_gy"], blob[tag + "_gx"], blob[tag + "_gw"] = \
            x.numpy(), w.numpy(), gy.numpy(), gr[0].numpy(), gr[1].numpy()

    x = randn(2, 2, 6000)
    w512 = torch.hann_window(512) + 0.05 * randn(512)
    record("stft512", x, w512, lambda x, w: ref.stft(x, 512, 128, window=w), lambda x, w: oc.stft(x, 512, 128, window=w))
    w200 = torch.hann_window(200) + 0.05 * randn(200)
    record("stft256_win200", x, w200, lambda x, w: ref.stft(x, 256, 64, win_length=200, window=w, normalized=True),
           lambda x, w: oc.stft(x, 256, 64, win_length=200, window=w, normalized=True))
    record("stft512_nocenter_two", x, w512, lambda x, w: ref.stft(x, 512, 200, window=w, center=False, onesided=False),
           lambda x, w: oc.stft(x, 512, 200, window=w, center=False, onesided=False))
    record("spec512_p1", x, w512, lambda x, w: ref.complex_norm(ref.stft(x, 512, 128, window=w), 1.0),
           lambda x, w: oc.complex_norm(oc.stft(x, 512, 128, window=w), 1.0))
    w2048 = torch.hann_window(2048) + 0.05 * randn(2048)
    x2 = randn(2, 1, 12000)
    fb = oc.mel_filterbank_for(64, 16000, fft_length=2048)
    record("meldb2048", x2, w2048,
           lambda x, w: ref.amplitude_to_db(ref.apply_filterbank(ref.complex_norm(ref.stft(x, 2048, 512, window=w), 2.0), fb)),
           lambda x, w: oc.amplitude_to_db(oc.apply_filterbank(oc.complex_norm(oc.stft(x, 2048, 512, window=w), 2.0), fb)))
    _save("grads_win.npz", **blob)
    print("window-gradient fixtures reproduce under oracle.ref_chain autograd")


def hpss_fixtures(ref, oc):
    """Harmonic-percussive separation (beta_hpss.py) from the UNMODIFIED reference module: soft and hard masks, powers 1 / 2 /
    0.7, kernel sizes 5 / 17 / 31, an input with a NaN.  `python oracle/gen_golden.py hpss` writes tests/golden/hpss.npz only."""
    import importlib
    ref_hpss = importlib.import_module("torchaudio_contrib.beta_hpss")
    g = torch.Generator().manual_seed(20261020)
    blob = {}
    x = torch.rand(2, 2, 40, 50, generator=g) ** 2 * 3.0
    blob["x"] = x.numpy()
    for tag, kw in (("k31_p2", dict(kernel_size=31, power=2.0)), ("k17_p1", dict(kernel_size=17, power=1.0)),
                    ("k5_p07", dict(kernel_size=5, power=0.7)), ("k31_hard", dict(kernel_size=31, power=2.0, hard=True)),
                    ("k9_maskonly", dict(kernel_size=(9, 9), power=2.0, mask_only=True))):
        r = ref_hpss.hpss(x, **kw)
        o = oc.hpss(x, **kw)
        for i, (a, b) in enumerate(zip(r, o)):
            if a is None:
                assert b is None
                continue
            _same(a, b, "hpss/%s[%d]" % (tag, i))
            blob["%s_%d" % (tag, i)] = a.numpy()
    xn = x.clone()
    xn[0, 1, 7, 9] = float("nan")
    r = ref_hpss.hpss(xn, 5, 1.0)
    o = oc.hpss(xn, 5, 1.0)
    for i, (a, b) in enumerate(zip(r, o)):
        assert torch.equal(torch.isnan(a), torch.isnan(b)) and torch.equal(torch.nan_to_num(a), torch.nan_to_num(b))
        blob["nan_%d" % i] = a.numpy()
    blob["xn"] = xn.numpy()
    assert repr(ref_hpss.HPSS(7, 1.0, True, False)) == "HPSS(kernel_size=7, power=1.0, hard=True, mask_only=False)"
    _save("hpss.npz", **blob)
    print("hpss fixtures reproduce under oracle.ref_chain")


def main():
    sys.path.insert(0, ROOT)
    from oracle import ref_chain as oc
    ref = _load_reference()
    if len(sys.argv) > 1 and sys.argv[1] == "widen":
        widen(ref, oc)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "grads":
        grads(ref, oc)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "grads_more":
        grads_more(ref, oc)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "hpss":
        hpss_fixtures(ref, oc)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "grads_win":
        grads_win(ref, oc)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "grads_pv":
        grads_pv(ref, oc)
        return
    if len(sys.argv) > 1:
        raise SystemExit("usage: gen_golden.py [widen | grads | grads_more | grads_pv | grads_win | hpss]   (no argument: the forward fixtures)")
    g = torch.Generator().manual_seed(20260925)

    def randn(*shape):
        return torch.randn(*shape, generator=g)

    # ---- config 1: Spectrogram(fft_length=512, hop_length=128) on (1,1,16000) -----------
    x = randn(1, 1, 16000)
    with _shimmed(ref):
        y = ref.Spectrogram(fft_length=512, hop_length=128)(x)
    _same(y, oc.spectrogram(x, 512, 128), "config 1")
    _save("cfg1_spectrogram_512_128.npz", x=x.numpy(), out=y.numpy())

    # ---- the reference's own STFT test configuration (fft 512 / hop 256, 2 channels) ------
    x = randn(1, 2, 6000)
    win = torch.hann_window(512)
    with _shimmed(ref):
        z = ref.stft(x, fft_length=512, hop_length=256, window=win, pad_mode='reflect')
    _same(z, oc.stft(x, 512, 256, window=win), "stft 512/256")
    _save("stft_512_256.npz", x=x.numpy(), out=z.contiguous().numpy())

    # ---- STFT option surface (H5): short window, normalized, no centring, pad modes, two-sided
    x = randn(3, 3000)
    cases = {
        "winlen": dict(fft_length=256, hop_length=64, win_length=200),
        "normalized": dict(fft_length=256, hop_length=100, normalized=True),
        "nocenter": dict(fft_length=512, hop_length=128, center=False),
        "constant": dict(fft_length=256, hop_length=64, pad_mode='constant'),
        "replicate": dict(fft_length=256, hop_length=64, pad_mode='replicate'),
        "circular": dict(fft_length=256, hop_length=64, pad_mode='circular'),
        "twosided": dict(fft_length=128, hop_length=32, onesided=False),
        "defaulthop": dict(fft_length=1024),
    }
    blob = {"x": x.numpy()}
    for tag, kw in cases.items():
        with _shimmed(ref):
            z = ref.STFT(**kw)(x)
        _same(z, oc.stft(x, **kw), "stft option " + tag)
        blob["out_" + tag] = z.contiguous().numpy()
    _save("stft_options.npz", **blob)

    # ---- config 2 shape family: Melspectrogram(128, 16 kHz, 2048/512), small batch ---------
    x = randn(2, 1, 16000)
    with _shimmed(ref):
        m = ref.Melspectrogram(num_mels=128, sample_rate=16000, fft_length=2048, hop_length=512)
        y = m(x)
        fb16 = m[2].filterbank
    _same(y, oc.melspectrogram(x, 128, 16000, fft_length=2048, hop_length=512), "config 2 (small)")
    _same(fb16, oc.mel_filterbank_for(128, 16000, fft_length=2048), "fb 16k")
    _save("mel_16k_2048_512.npz", x=x.numpy(), out=y.contiguous().numpy())

    # ---- config 3 family: 48 kHz, 2 channels, + AmplitudeToDb ------------------------------
    x = randn(1, 2, 24000)
    with _shimmed(ref):
        m = torch.nn.Sequential(*ref.Melspectrogram(num_mels=128, sample_rate=48000,
                                                    fft_length=2048, hop_length=512),
                                ref.AmplitudeToDb())
        y = m(x)
        fb48 = m[2].filterbank
    _same(y, oc.melspectrogram(x, 128, 48000, to_db=True, fft_length=2048, hop_length=512), "config 3 (small)")
    _save("meldb_48k_2048_512.npz", x=x.numpy(), out=y.contiguous().numpy())

    # ---- filterbank matrices (pinned by no reference test; pinned here by the reference run) --
    fbs = {"fb_16k_1025x128": fb16, "fb_48k_1025x128": fb48}
    for fft in (256, 512, 1024, 4096):
        with _shimmed(ref):
            f = ref.MelFilterbank(num_freqs=fft // 2 + 1, num_mels=128, sample_rate=16000).get_filterbank()
        _same(f, oc.mel_filterbank_for(128, 16000, fft_length=fft), "fb fft %d" % fft)
        fbs["fb_16k_%dx128" % (fft // 2 + 1)] = f
    with _shimmed(ref):
        f = ref.MelFilterbank(num_freqs=1025, num_mels=40, sample_rate=22050, htk=True, min_freq=30.0).get_filterbank()
    _same(f, oc.create_mel_filter(1025, 40, 30.0, 22050 // 2, True), "fb htk")
    fbs["fb_htk_22k_1025x40"] = f
    with _shimmed(ref):
        f = ref.MelFilterbank(num_freqs=257, num_mels=128, max_freq=1.0).get_filterbank()   # tests/test_layers.py:97
    _same(f, oc.create_mel_filter(257, 128, 0.0, 1.0, False), "fb max_freq=1")
    fbs["fb_maxfreq1_257x128"] = f
    _save("filterbanks.npz", **{k: v.numpy() for k, v in fbs.items()})

    # ---- fft sweep (config 5b): fft 256..4096, hop = fft/4, 128 mels, 16 kHz ---------------
    x = randn(2, 1, 12000)
    blob = {"x": x.numpy()}
    for fft in (256, 512, 1024, 2048, 4096):
        with _shimmed(ref):
            y = ref.Melspectrogram(num_mels=128, sample_rate=16000, fft_length=fft, hop_length=fft // 4)(x)
        _same(y, oc.melspectrogram(x, 128, 16000, fft_length=fft, hop_length=fft // 4), "sweep %d" % fft)
        blob["out_%d" % fft] = y.contiguous().numpy()
    _save("mel_sweep_16k.npz", **blob)

    # ---- complex_norm / apply_filterbank / amplitude_to_db as separate stages --------------
    z = randn(2, 257, 50, 2)
    fbr = randn(257, 36)                                     # dense random matrix (tests/test_functional.py:137)
    p07 = ref.complex_norm(z, 0.7)
    _same(p07, oc.complex_norm(z, 0.7), "complex_norm 0.7")
    mag = ref.complex_norm(z, 1.0)
    afb = ref.apply_filterbank(mag, fbr)
    _same(afb, oc.apply_filterbank(mag, fbr), "apply_filterbank")
    db = ref.amplitude_to_db(mag, ref=2.0, amin=1e-5)
    _same(db, oc.amplitude_to_db(mag, 2.0, 1e-5), "amplitude_to_db")
    _save("stages.npz", z=z.numpy(), fb=fbr.numpy(), norm_p07=p07.numpy(), norm_p1=mag.numpy(),
          filtered=afb.contiguous().numpy(), db=db.numpy())

    # ---- mu-law (config 5a) ----------------------------------------------------------------
    xu = torch.rand(4096, generator=g) * 2 - 1                # in range
    xo = 2 * (randn(4096) - 0.5)                              # the reference test's distribution, mostly out of range
    edge = torch.tensor([0.0, -0.0, 1.0, -1.0, 0.5, -0.5, 1e-8, -1e-8, 3.0, -3.0, 1e30, -1e30,
                         1.1754944e-38, 65504.0, 0.99999994, -0.99999994])
    xs = torch.cat([xu, xo, edge])
    enc = ref.mu_law_encoding(xs, 256)
    _same(enc, oc.mu_law_encoding(xs, 256), "mu-law encode")
    codes = torch.arange(256)
    dec = ref.mu_law_decoding(codes, 256)
    _same(dec, oc.mu_law_decoding(codes, 256), "mu-law decode")
    rt = ref.mu_law_encoding(dec, 256)
    enc64 = ref.mu_law_encoding(xs, 64)
    dec64 = ref.mu_law_decoding(torch.arange(64), 64)
    _same(enc64, oc.mu_law_encoding(xs, 64), "mu-law encode q64")
    _save("mulaw.npz", x=xs.numpy(), enc256=enc.numpy(), dec256=dec.numpy(), roundtrip256=rt.numpy(),
          enc64=enc64.numpy(), dec64=dec64.numpy())
    print("all fixtures reproduce under oracle.ref_chain")
    widen(ref, oc)


if __name__ == "__main__":
    main()
