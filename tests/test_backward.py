"""Backward passes (csrc/stft_backward.cu): d loss / d input against the committed gradients of the
UNMODIFIED reference under torch autograd (tests/golden/grads.npz, oracle/gen_golden.py grads) and against the oracle's
autograd on fresh inputs.  Tolerance: 1e-4 of the gradient's rms level (north_star: 1e-4 relative, fp32)."""
import pytest
import torch

from conftest import golden, rel_err
from oracle import ref_chain as oc

REL = 1e-4

STFT_CASES = {
    "stft_512_128": dict(fft_length=512, hop_length=128),
    "stft_winlen_norm": dict(fft_length=256, hop_length=64, win_length=200, normalized=True),
    "stft_nocenter": dict(fft_length=512, hop_length=100, center=False),
    "stft_constant": dict(fft_length=256, hop_length=64, pad_mode='constant'),
    "stft_replicate": dict(fft_length=256, hop_length=64, pad_mode='replicate'),
    "stft_circular": dict(fft_length=256, hop_length=64, pad_mode='circular'),
    "stft_twosided": dict(fft_length=128, hop_length=32, onesided=False),
    "stft_2048": dict(fft_length=2048, hop_length=512),
}


def _oracle_grad(fn, x, gy):
    xo = x.clone().requires_grad_(True)
    (gx,) = torch.autograd.grad(fn(xo), xo, gy)
    return gx


# --------------------------------------------------------------------------------------------- CPU: oracle vs fixtures
def test_oracle_autograd_reproduces_reference_gradients():
    g = golden("grads.npz")
    for tag, kw in STFT_CASES.items():
        win = torch.hann_window(kw.get("win_length", kw["fft_length"]))
        gx = _oracle_grad(lambda t: oc.stft(t, window=win, **kw), g[tag + "_x"], g[tag + "_gy"])
        assert rel_err(gx, g[tag + "_gx"]) < 1e-5, tag
    gx = _oracle_grad(lambda t: oc.melspectrogram(t, 128, 48000, to_db=True, fft_length=2048, hop_length=512),
                      g["mel_2048_db_x"], g["mel_2048_db_gy"])
    assert rel_err(gx, g["mel_2048_db_gx"]) < 1e-5
    gx = _oracle_grad(lambda t: oc.amplitude_to_db(t, 1.0, 1e-7), g["todb_x"], g["todb_gy"])
    assert torch.allclose(gx, g["todb_gx"], rtol=1e-6, atol=0)


def test_requires_grad_on_constants_is_refused_without_a_gpu_call():
    """Where an argument is a constant of the path taken (the window of the float64 path, phase_advance), requiring a gradient
    for it raises before any library call -- never a silent detach.  (The float32 window IS differentiated: test_window_gradient.)"""
    import torchaudio_contrib_b200.functional as F
    w = torch.hann_window(512).requires_grad_(True)
    with pytest.raises(RuntimeError, match="is a constant on this path"):
        F._no_param_grad(w, "stft: window")
    x = torch.zeros(1, 1, 4096, dtype=torch.float64)    # the check comes first: CPU tensor, no library call
    for fn in (lambda: F.stft(x, 512, window=w), lambda: F.spectrogram(x, 512, window=w)):
        with pytest.raises(RuntimeError, match="is a constant on this path"):
            fn()
    with pytest.raises(RuntimeError, match="is a constant on this path"):
        F.phase_vocoder(torch.zeros(1, 5, 8, 2), 1.3, torch.zeros(5, 1, requires_grad=True))


def test_oracle_autograd_reproduces_round2_gradients():
    """tests/golden/grads_more.npz (oracle/gen_golden.py grads_more): pointwise operators and the filterbank gradient."""
    g = golden("grads_more.npz")

    def grads(fn, names, tag, n_out=1):
        leaves = [g["%s_in_%s" % (tag, k)].clone().requires_grad_(True) for k in names]
        out = fn(*leaves)
        out = out if isinstance(out, tuple) else (out,)
        return torch.autograd.grad(out, leaves, [g["%s_gy%d" % (tag, i)] for i in range(n_out)])

    cases = [("fromdb", ["x"], lambda x: oc.db_to_amplitude(x, 2.0), 1), ("angle", ["z"], oc.angle, 1),
             ("magphase_p1", ["z"], lambda z: oc.magphase(z, 1.0), 2), ("magphase_p2", ["z"], lambda z: oc.magphase(z, 2.0), 2),
             ("mudec", ["c"], lambda c: oc.mu_law_decoding(c, 256), 1),
             ("fbank_param", ["s", "fb"], oc.apply_filterbank, 1)]
    for tag, names, fn, n_out in cases:
        for k, got in zip(names, grads(fn, names, tag, n_out)):
            assert rel_err(got, g["%s_g_%s" % (tag, k)]) < 1e-5, (tag, k)


# --------------------------------------------------------------------------------------------- GPU: kernels vs fixtures
@pytest.fixture(scope="module")
def tac():
    import torchaudio_contrib_b200 as tac
    return tac


def _gpu_grad(fn, x, gy):
    xg = x.cuda().requires_grad_(True)
    y = fn(xg)
    assert y.requires_grad
    (gx,) = torch.autograd.grad(y, xg, gy.cuda())
    return y.detach().cpu(), gx.cpu()


@pytest.mark.gpu
@pytest.mark.parametrize("tag", sorted(STFT_CASES))
def test_stft_backward(tac, tag):
    g = golden("grads.npz")
    kw = STFT_CASES[tag]
    win = torch.hann_window(kw.get("win_length", kw["fft_length"])).cuda()
    _, gx = _gpu_grad(lambda t: tac.stft(t, window=win, **kw), g[tag + "_x"], g[tag + "_gy"])
    assert gx.shape == g[tag + "_gx"].shape
    assert rel_err(gx, g[tag + "_gx"]) < REL


@pytest.mark.gpu
@pytest.mark.parametrize("tag,power", [("spec_p1", 1.0), ("spec_p2", 2.0), ("spec_p0_7", 0.7)])
def test_spectrogram_backward(tac, tag, power):
    g = golden("grads.npz")
    model = tac.Spectrogram(fft_length=512, hop_length=128, power=power).cuda()
    _, gx = _gpu_grad(model, g[tag + "_x"], g[tag + "_gy"])
    assert rel_err(gx, g[tag + "_gx"]) < REL


@pytest.mark.gpu
def test_spectrogram_twosided_backward(tac):
    g = golden("grads.npz")
    model = tac.Spectrogram(fft_length=128, hop_length=32, onesided=False, power=2.0).cuda()      # child-by-child path
    _, gx = _gpu_grad(model, g["spec_twosided_x"], g["spec_twosided_gy"])
    assert rel_err(gx, g["spec_twosided_gx"]) < REL
    _, gx = _gpu_grad(lambda t: tac.functional.spectrogram(t, 128, 32, onesided=False, power=2.0),
                      g["spec_twosided_x"], g["spec_twosided_gy"])
    assert rel_err(gx, g["spec_twosided_gx"]) < REL


@pytest.mark.gpu
def test_melspectrogram_backward(tac):
    g = golden("grads.npz")
    model = tac.Melspectrogram(num_mels=128, sample_rate=16000, fft_length=2048, hop_length=512).cuda()     # one-kernel forward
    y, gx = _gpu_grad(model, g["mel_2048_x"], g["mel_2048_gy"])
    assert y.shape == g["mel_2048_gy"].shape
    assert rel_err(gx, g["mel_2048_gx"]) < REL
    model = tac.Sequential(*tac.Melspectrogram(num_mels=128, sample_rate=48000, fft_length=2048, hop_length=512),
                           tac.AmplitudeToDb()).cuda()
    _, gx = _gpu_grad(model, g["mel_2048_db_x"], g["mel_2048_db_gy"])
    assert rel_err(gx, g["mel_2048_db_gx"]) < REL
    model = tac.Melspectrogram(num_mels=64, sample_rate=22050, fft_length=1024, hop_length=256).cuda()      # two-kernel forward
    _, gx = _gpu_grad(model, g["mel_1024_64_x"], g["mel_1024_64_gy"])
    assert rel_err(gx, g["mel_1024_64_gx"]) < REL


@pytest.mark.gpu
def test_unfused_chain_backward_matches_fused(tac):
    """nn.Sequential of the same children (every module on its own backward kernel) against the oracle's autograd."""
    torch.manual_seed(5)
    x = torch.randn(3, 1, 9000)
    mods = list(tac.Melspectrogram(num_mels=40, sample_rate=16000, fft_length=512, hop_length=160)) + [tac.AmplitudeToDb()]
    plain = torch.nn.Sequential(*mods).cuda()
    fused = tac.Sequential(*mods).cuda()
    y0 = oc.melspectrogram(x, 40, 16000, to_db=True, fft_length=512, hop_length=160)
    gy = torch.randn(y0.shape)
    want = _oracle_grad(lambda t: oc.melspectrogram(t, 40, 16000, to_db=True, fft_length=512, hop_length=160), x, gy)
    _, g_plain = _gpu_grad(plain, x, gy)
    _, g_fused = _gpu_grad(fused, x, gy)
    assert rel_err(g_plain, want) < REL
    assert rel_err(g_fused, want) < REL


@pytest.mark.gpu
@pytest.mark.parametrize("tag,power", [("cnorm_p1", 1.0), ("cnorm_p2", 2.0), ("cnorm_p0_5", 0.5)])
def test_complex_norm_backward(tac, tag, power):
    g = golden("grads.npz")
    _, gx = _gpu_grad(lambda t: tac.complex_norm(t, power), g[tag + "_x"], g[tag + "_gy"])
    assert torch.isfinite(gx).all()
    assert rel_err(gx, g[tag + "_gx"]) < 1e-5


@pytest.mark.gpu
def test_apply_filterbank_and_db_backward(tac):
    g = golden("grads.npz")
    fb = g["fb_dense"].cuda()
    _, gx = _gpu_grad(lambda t: tac.apply_filterbank(t, fb), g["fbank_dense_x"], g["fbank_dense_gy"])
    assert rel_err(gx, g["fbank_dense_gx"]) < 1e-5
    # gradient arriving as a transposed view (what the reference-layout output produces downstream)
    gy_t = g["fbank_dense_gy"].transpose(-2, -1).contiguous().transpose(-2, -1)
    _, gx = _gpu_grad(lambda t: tac.apply_filterbank(t, fb), g["fbank_dense_x"], gy_t)
    assert rel_err(gx, g["fbank_dense_gx"]) < 1e-5
    _, gx = _gpu_grad(lambda t: tac.amplitude_to_db(t, 1.0, 1e-7), g["todb_x"], g["todb_gy"])
    want = g["todb_gx"]
    assert torch.equal(gx == 0, want == 0)                           # the amin clamp switches the gradient off at the same inputs
    assert torch.allclose(gx, want, rtol=1e-5, atol=0)


@pytest.mark.gpu
def test_backward_large_shape_linearity(tac):
    """BASELINE config 2 shape: the adjoint is linear in the upstream gradient and <grad_x, dx> = <gy, J dx> (the
    defining property of the adjoint, checked with the forward kernel as J for the quadratic-free stft)."""
    torch.manual_seed(9)
    x = torch.randn(8, 1, 160000, device="cuda")
    dx = torch.randn_like(x)
    xg = x.clone().requires_grad_(True)
    y = tac.stft(xg, 2048, 512)
    gy = torch.randn_like(y)
    (gx,) = torch.autograd.grad(y, xg, gy)
    jdx = tac.stft(dx, 2048, 512)                                     # stft is linear: J dx = stft(dx)
    lhs = (gx.double() * dx.double()).sum().item()
    rhs = (gy.double() * jdx.double()).sum().item()
    assert abs(lhs - rhs) <= 1e-5 * gx.double().norm().item() * dx.double().norm().item()
    (gx2,) = torch.autograd.grad(tac.stft(xg, 2048, 512), xg, 2.0 * gy)
    assert rel_err(gx2.cpu(), (2.0 * gx).cpu()) < 1e-5


@pytest.mark.gpu
def test_no_grad_paths_still_refuse_silent_detach(tac):
    """What has no adjoint kernel raises instead of silently detaching: mu_law_encoding (integer output), phase_advance,
    double inputs of the signal path."""
    x = torch.randn(2, 1, 4000, device="cuda", requires_grad=True)
    z = tac.stft(x.detach(), 512, 128)
    adv = torch.linspace(0, 3.14159 * 128, 257, device="cuda")[..., None]
    with pytest.raises(RuntimeError):
        tac.phase_vocoder(z, 1.3, adv.requires_grad_(True))
    with pytest.raises(RuntimeError):
        tac.mu_law_encoding(x)
    with pytest.raises(NotImplementedError):
        tac.stft(x.double(), 512, 128)
    tac.stft(x, 400, 160).sum().backward()                  # sizes that are not a power of two differentiate (direct-DFT adjoint)
    assert x.grad is not None and bool(torch.isfinite(x.grad).all())


@pytest.mark.gpu
@pytest.mark.parametrize("power", [2.0, 1.0, 0.7])
@pytest.mark.parametrize("n_samples,pad_mode,normalized", [(2049, "reflect", False), (9000, "constant", True),
                                                            (16001, "replicate", False), (12288, "circular", False)])
def test_mel_2048_backward_options(tac, power, n_samples, pad_mode, normalized):
    """The warp-per-frame adjoint kernel (n_fft 2048) against the oracle's autograd: exponents, padding modes, ragged
    lengths (gather path), fewer frames than warps, a frame-major and a contiguous upstream gradient."""
    torch.manual_seed(n_samples + int(10 * power))
    x = torch.randn(2, 1, n_samples)
    fb = tac.MelFilterbank(num_freqs=1025, num_mels=128, sample_rate=16000).get_filterbank()

    def ref(t):
        spec = oc.spectrogram(t, 2048, 512, pad_mode=pad_mode, normalized=normalized, power=power)
        return oc.apply_filterbank(spec, fb)

    y0 = ref(x)
    gy = torch.randn(y0.shape)
    want = _oracle_grad(ref, x, gy)
    # exponents below 1 make the gradient singular at small |X| (|X|^(p-2)): two float32 evaluations then differ by more
    # than 1e-4 from each other, so both are judged against the float64 evaluation of the same chain
    win64 = torch.hann_window(2048, dtype=torch.float64)

    def ref64(t):
        spec = oc.spectrogram(t, 2048, 512, window=win64, pad_mode=pad_mode, normalized=normalized, power=power)
        return oc.apply_filterbank(spec, fb.double())

    want64 = _oracle_grad(ref64, x.double(), gy.double())
    tol = max(REL, 10.0 * rel_err(want, want64))            # p = 0.7: the fp32 reference itself is ~6e-5 off
    for layout in ("reference", "contiguous"):
        _, gx = _gpu_grad(lambda t: tac.functional.melspectrogram(t, fb.cuda(), 2048, 512, pad_mode=pad_mode, normalized=normalized,
                                                                  power=power, layout=layout), x, gy)
        assert rel_err(gx, want64) < tol, (layout, tol)


@pytest.mark.gpu
def test_backward_is_deterministic(tac):
    """No atomics on the backward path: the same input and upstream gradient give the same bits."""
    torch.manual_seed(2)
    x = torch.randn(4, 1, 40000)
    gy = None
    outs = []
    m = tac.Sequential(*tac.Melspectrogram(num_mels=128, sample_rate=16000, fft_length=2048, hop_length=512), tac.AmplitudeToDb()).cuda()
    for _ in range(3):
        xg = x.cuda().requires_grad_(True)
        y = m(xg)
        if gy is None:
            gy = torch.randn_like(y)
        (gx,) = torch.autograd.grad(y, xg, gy)
        outs.append(gx)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])


@pytest.mark.gpu
@pytest.mark.parametrize("power,n_samples,pad_mode", [(2.0, 20000, "reflect"), (1.0, 7001, "reflect"), (2.0, 4096, "constant"),
                                                      (0.7, 12288, "replicate")])
def test_spectrogram_2048_backward(tac, power, n_samples, pad_mode):
    """Spectrogram at n_fft 2048: the public-layout gradient is transposed to frame-major and takes the warp-per-frame
    adjoint kernel."""
    torch.manual_seed(int(n_samples + 10 * power))
    x = torch.randn(3, 2, n_samples)
    win64 = torch.hann_window(2048, dtype=torch.float64)

    def ref(t):
        return oc.spectrogram(t, 2048, 512, pad_mode=pad_mode, power=power)

    gy = torch.randn(ref(x).shape)
    want = _oracle_grad(ref, x, gy)
    want64 = _oracle_grad(lambda t: oc.spectrogram(t, 2048, 512, window=win64, pad_mode=pad_mode, power=power), x.double(), gy.double())
    # p <= 1: d|X|^p / dX ~ X / |X|^(2-p) is ill-conditioned at bins that vanish in exact arithmetic (the reflected edge
    # frames are symmetric, so whole sets of bins do): any fp32 evaluation is off by O(rounding / |X|) there
    tol = max(REL if power > 1.0 else 3e-4, 10.0 * rel_err(want, want64))
    model = tac.Spectrogram(fft_length=2048, hop_length=512, pad_mode=pad_mode, power=power).cuda()
    _, gx = _gpu_grad(model, x, gy)
    assert rel_err(gx, want64) < tol


@pytest.mark.gpu
def test_mel_2048_backward_dense_filterbank(tac):
    """A dense (random) filterbank at n_fft 2048: forward on the tensor-core path, backward through the wide-row branch
    of the frame-major filterbank adjoint and the warp-per-frame kernel."""
    torch.manual_seed(8)
    x = torch.randn(2, 1, 14000)
    fb = torch.randn(1025, 24).abs()

    def ref(t):
        return oc.apply_filterbank(oc.spectrogram(t, 2048, 512, power=2.0), fb)

    gy = torch.randn(ref(x).shape)
    want = _oracle_grad(ref, x, gy)
    y, gx = _gpu_grad(lambda t: tac.functional.melspectrogram(t, fb.cuda(), 2048, 512), x, gy)
    assert rel_err(y, ref(x)) < REL
    assert rel_err(gx, want) < REL


# --------------------------------------------------------------------------------------------- round 2: N4 pointwise + filterbank
@pytest.mark.gpu
def test_pointwise_gradients_round2(tac):
    """db_to_amplitude, angle, magphase, float mu_law_decoding: gradients of the unmodified reference (grads_more.npz)."""
    g = golden("grads_more.npz")

    def check(tag, fn, n_out=1, tol=REL):
        name = [k for k in g if k.startswith(tag + "_in_")][0].split("_in_")[1]
        leaf = g["%s_in_%s" % (tag, name)].cuda().requires_grad_(True)
        out = fn(leaf)
        out = out if isinstance(out, tuple) else (out,)
        (gx,) = torch.autograd.grad(out, [leaf], [g["%s_gy%d" % (tag, i)].cuda() for i in range(n_out)])
        assert rel_err(gx.cpu(), g["%s_g_%s" % (tag, name)]) < tol, tag

    check("fromdb", lambda x: tac.db_to_amplitude(x, ref=2.0))
    check("angle", tac.angle)
    check("magphase_p1", lambda z: tac.magphase(z, 1.0), 2)
    check("magphase_p2", lambda z: tac.magphase(z, 2.0), 2)
    check("mudec", lambda c: tac.mu_law_decoding(c, 256))
    # only one of magphase's outputs used: the other gradient is absent, not zero-filled garbage
    z = g["angle_in_z"].cuda().requires_grad_(True)
    mag, _ = tac.magphase(z, 1.0)
    (gz,) = torch.autograd.grad(mag.sum(), z)
    zr = g["angle_in_z"].clone().requires_grad_(True)
    (want,) = torch.autograd.grad(oc.magphase(zr, 1.0)[0].sum(), zr)
    assert rel_err(gz.cpu(), want) < REL


@pytest.mark.gpu
def test_filterbank_parameter_gradient(tac):
    """A learnable filterbank (requires_grad) gets the reference's gradient -- with or without a gradient for the signal,
    through apply_filterbank and through the Melspectrogram chain (+ dB), one-kernel and two-kernel paths."""
    g = golden("grads_more.npz")
    s = g["fbank_param_in_s"].cuda().requires_grad_(True)
    fb = g["fbank_param_in_fb"].cuda().requires_grad_(True)
    gy = g["fbank_param_gy0"].cuda()
    gs, gfb = torch.autograd.grad(tac.apply_filterbank(s, fb), [s, fb], gy)
    assert rel_err(gs.cpu(), g["fbank_param_g_s"]) < REL and rel_err(gfb.cpu(), g["fbank_param_g_fb"]) < REL
    (gfb_only,) = torch.autograd.grad(tac.apply_filterbank(s.detach(), fb), [fb], gy)       # signal without grad
    assert rel_err(gfb_only.cpu(), g["fbank_param_g_fb"]) < REL

    x = g["melchain_param_in_x"].cuda().requires_grad_(True)
    fb = g["melchain_param_in_fb"].cuda().requires_grad_(True)
    gy = g["melchain_param_gy0"].cuda()
    y = tac.functional.melspectrogram(x, fb, 2048, 512, to_db=True)
    gx, gfb = torch.autograd.grad(y, [x, fb], gy)
    assert rel_err(gx.cpu(), g["melchain_param_g_x"]) < REL
    assert rel_err(gfb.cpu(), g["melchain_param_g_fb"]) < REL
    y = tac.functional.melspectrogram(x.detach(), fb, 2048, 512, to_db=True)                # parameter only
    assert y.requires_grad
    (gfb2,) = torch.autograd.grad(y, [fb], gy)
    assert rel_err(gfb2.cpu(), g["melchain_param_g_fb"]) < REL


# --------------------------------------------------------------------------------------------- round 2: phase vocoder
def test_oracle_autograd_reproduces_phase_vocoder_gradients():
    """tests/golden/grads_pv.npz (oracle/gen_golden.py grads_pv): the unmodified reference's float64 gradients."""
    g = golden("grads_pv.npz")
    prior = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        for tag in ("r07", "r13", "r20", "r03"):
            z = g[tag + "_z"].clone().requires_grad_(True)
            y = oc.phase_vocoder(z, float(g[tag + "_rate"]), g[tag + "_adv"])
            (gz,) = torch.autograd.grad(y, z, g[tag + "_gy"])
            assert torch.allclose(gz, g[tag + "_gz"], rtol=1e-10, atol=1e-10), tag
    finally:
        torch.set_default_dtype(prior)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["r07", "r13", "r20", "r03"])
def test_phase_vocoder_backward(tac, tag):
    """functional.py:204-274 under autograd: the gather kernel (csrc/phase_vocoder.cu) against the reference's float64
    gradient -- float64 tensors to 1e-9 of the gradient's scale, float32 tensors (float64 inside) to float32 rounding;
    deterministic; the TimeStretch module differentiates too."""
    g = golden("grads_pv.npz")
    rate = float(g[tag + "_rate"])
    want = g[tag + "_gz"]
    scale = want.abs().max().item()
    prior = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)                      # the reference's time steps are built in the default dtype
    try:
        z = g[tag + "_z"].cuda().requires_grad_(True)
        y = tac.phase_vocoder(z, rate, g[tag + "_adv"].cuda())
        assert (y.detach().cpu() - g[tag + "_y"]).abs().max().item() < 1e-8
        (gz,) = torch.autograd.grad(y, z, g[tag + "_gy"].cuda())
        assert gz.dtype == torch.float64 and gz.shape == want.shape
        assert (gz.cpu() - want).abs().max().item() < 1e-9 * scale, tag
        z2 = g[tag + "_z"].cuda().requires_grad_(True)
        (again,) = torch.autograd.grad(tac.phase_vocoder(z2, rate, g[tag + "_adv"].cuda()), z2, g[tag + "_gy"].cuda())
        assert torch.equal(gz, again)
        z32 = g[tag + "_z"].float().cuda().requires_grad_(True)
        (g32,) = torch.autograd.grad(tac.phase_vocoder(z32, rate, g[tag + "_adv"].float().cuda()), z32, g[tag + "_gy"].float().cuda())
        assert g32.dtype == torch.float32 and (g32.cpu().double() - want).abs().max().item() < 2e-4 * scale, tag
    finally:
        torch.set_default_dtype(prior)
    with pytest.raises(RuntimeError):
        tac.phase_vocoder(g[tag + "_z"].cuda(), rate, g[tag + "_adv"].cuda().requires_grad_(True))


# --------------------------------------------------------------------------------------------- round 2: learnable window
@pytest.mark.gpu
def test_window_gradient(tac):
    """A window that requires grad gets the reference's gradient (tests/golden/grads_win.npz, from the unmodified reference):
    complex stft with a full-length and a shorter (centre-padded) window, normalized, two-sided without centring; the
    Spectrogram and mel + dB chains (functional and modules), with and without a gradient for the waveform."""
    g = golden("grads_win.npz")
    F = tac.functional

    def run(tag, fn, want_x=True):
        x = g[tag + "_x"].cuda().requires_grad_(want_x)
        w = g[tag + "_w"].cuda().requires_grad_(True)
        y = fn(x, w)
        assert y.requires_grad
        grads = torch.autograd.grad(y, [x, w] if want_x else [w], g[tag + "_gy"].cuda())
        gw = grads[-1]
        assert gw.shape == g[tag + "_gw"].shape
        assert rel_err(gw.cpu(), g[tag + "_gw"]) < REL, (tag, "window", rel_err(gw.cpu(), g[tag + "_gw"]))
        if want_x:
            assert rel_err(grads[0].cpu(), g[tag + "_gx"]) < REL, (tag, "waveform")

    for want_x in (True, False):
        run("stft512", lambda x, w: F.stft(x, 512, 128, window=w), want_x)
        run("stft256_win200", lambda x, w: F.stft(x, 256, 64, win_length=200, window=w, normalized=True), want_x)
        run("stft512_nocenter_two", lambda x, w: F.stft(x, 512, 200, window=w, center=False, onesided=False), want_x)
        run("spec512_p1", lambda x, w: F.spectrogram(x, 512, 128, window=w, power=1.0), want_x)
    fb = oc.mel_filterbank_for(64, 16000, fft_length=2048).cuda()
    run("meldb2048", lambda x, w: F.melspectrogram(x, fb, 2048, 512, window=w, to_db=True))
    # module: a window turned into a parameter
    mel = tac.Melspectrogram(num_mels=64, sample_rate=16000, fft_length=2048, hop_length=512).cuda()
    stft_mod = mel[0]
    del stft_mod._buffers["window"]
    stft_mod.window = torch.nn.Parameter(g["meldb2048_w"].cuda())
    x = g["meldb2048_x"].cuda()
    y = tac.AmplitudeToDb().cuda()(mel(x))
    (gw,) = torch.autograd.grad(y, [stft_mod.window], g["meldb2048_gy"].cuda())
    assert rel_err(gw.cpu(), g["meldb2048_gw"]) < REL
