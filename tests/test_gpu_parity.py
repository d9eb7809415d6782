"""Parity of the CUDA path (through the C ABI) against the oracle and the committed golden vectors.
Tolerances: BASELINE.json north star -- 1e-4 relative fp32 for the mel path (dB: 1e-3 dB absolute),
bit-exact for mu-law codes."""
import os

import numpy as np
import pytest
import torch

from conftest import golden, pure_rel_err, rel_err

pytestmark = pytest.mark.gpu

REL = 1e-4


@pytest.fixture(scope="module")
def tac():
    import torchaudio_contrib_b200 as t
    return t


@pytest.fixture(scope="module")
def oc():
    from oracle import ref_chain
    return ref_chain


def dev(t):
    return t.cuda()


# ------------------------------------------------------------------------------------------ mu-law
def test_mulaw_golden(tac):
    g = golden("mulaw.npz")
    x = dev(g["x"])
    assert torch.equal(tac.mu_law_encoding(x, 256).cpu(), g["enc256"])
    assert torch.equal(tac.mu_law_encoding(x, 64).cpu(), g["enc64"])
    codes = dev(torch.arange(256))
    dec = tac.mu_law_decoding(codes, 256)
    assert torch.equal(dec.cpu(), g["dec256"])
    assert torch.equal(tac.mu_law_encoding(dec, 256).cpu(), g["roundtrip256"])
    assert torch.equal(tac.mu_law_decoding(dev(torch.arange(64)), 64).cpu(), g["dec64"])


@pytest.mark.parametrize("shape", [(1, 100000), (1, 2, 100000), (7,), (3, 1, 1021)])
def test_mulaw_reference_test_distribution(tac, oc, shape):
    """tests/test_functional.py:161-203 restated: 2*(randn-0.5) is mostly out of [-1, 1]."""
    torch.manual_seed(7)
    w = 2 * (torch.randn(*shape) - 0.5)
    assert torch.equal(tac.mu_law_encoding(dev(w), 256).cpu(), oc.mu_law_encoding(w, 256))
    codes = torch.randint(0, 255, (1, 1024))
    ref = oc.mu_law_decoding(codes.float(), 256)
    assert torch.equal(tac.mu_law_decoding(dev(codes.float()), 256).cpu(), ref)
    assert torch.equal(tac.mu_law_decoding(dev(codes), 256).cpu(), ref)
    assert torch.equal(tac.mu_law_encoding(tac.mu_law_decoding(dev(codes), 256), 256).cpu(), codes)


def test_mulaw_modules_and_edges(tac, oc):
    enc, decm = tac.MuLawEncoding().cuda(), tac.MuLawDecoding().cuda()
    x = torch.tensor([0.0, -0.0, 1.0, -1.0, 3.0, -3.0, 1e30, -1e30, float("inf"), float("-inf"), float("nan"),
                      1e-40, 1.3344405e36, 1.3344407e36, 3.4e38])
    got = enc(dev(x)).cpu()
    want = oc.mu_law_encoding(x, 256)
    assert torch.equal(got, want)
    ints = torch.tensor([-3, 0, 5, 255, 256, 1000])
    out = decm(dev(ints)).cpu()
    ref = oc.mu_law_decoding(ints, 256)
    assert torch.equal(out[1:4], ref[1:4])                       # table range: exact
    assert pure_rel_err(out, ref) < 1e-5                          # closed form outside the table
    assert enc(dev(torch.arange(-5, 5, dtype=torch.int32))).dtype == torch.int64


@pytest.mark.parametrize("q", [2, 16, 1024, 65536])
def test_mulaw_other_levels(tac, oc, q):
    torch.manual_seed(q)
    x = torch.cat([torch.rand(200000) * 2 - 1, torch.randn(50000) * 3])
    assert torch.equal(tac.mu_law_encoding(dev(x), q).cpu(), oc.mu_law_encoding(x, q))
    if q <= 1024:
        codes = torch.arange(q)
        assert torch.equal(tac.mu_law_decoding(dev(codes), q).cpu(), oc.mu_law_decoding(codes, q))


@pytest.mark.slow
def test_mulaw_exhaustive_all_floats(tac, oc):
    """Every one of the 2^32 float bit patterns: CUDA codes == reference codes."""
    chunk = 1 << 27
    bad = 0
    for start in range(0, 1 << 32, chunk):
        bits = torch.arange(start, start + chunk, dtype=torch.int64)
        bits = torch.where(bits >= (1 << 31), bits - (1 << 32), bits).to(torch.int32)
        x = bits.view(torch.float32)
        got = tac.mu_law_encoding(dev(x), 256).cpu()
        want = oc.mu_law_encoding(x, 256)
        bad += int((got != want).sum())
    assert bad == 0


# ------------------------------------------------------------------------------------------ pointwise
def test_stages_golden(tac):
    g = golden("stages.npz")
    z = dev(g["z"])
    assert rel_err(tac.complex_norm(z, 0.7).cpu(), g["norm_p07"]) < REL
    mag = tac.complex_norm(z, 1.0)
    assert rel_err(mag.cpu(), g["norm_p1"]) < 1e-6
    out = tac.apply_filterbank(dev(g["norm_p1"]), dev(g["fb"])).cpu()
    assert out.shape == g["filtered"].shape
    assert rel_err(out, g["filtered"]) < REL
    db = tac.amplitude_to_db(dev(g["norm_p1"]), ref=2.0, amin=1e-5).cpu()
    assert (db - g["db"]).abs().max().item() < 1e-3


def test_amplitude_db_known_answers(tac):
    """tests/test_functional.py:144-158: power [1e-6..1e6] <-> dB [-60..60]."""
    power = torch.tensor([0.000001, 0.0001, 0.1, 1.0, 10.0, 1000000.0])
    db = torch.tensor([-60.0, -40.0, -10.0, 0.0, 10.0, 60.0])
    got = tac.amplitude_to_db(dev(power.sqrt()), ref=1.0).cpu()
    assert (got - db).abs().max().item() < 1e-5


@pytest.mark.parametrize("shape", [(1, 2, 1025, 400, 2), (1025, 400, 2)])
@pytest.mark.parametrize("power", [1, 2, 0.7])
def test_complex_norm(tac, shape, power):
    """tests/test_functional.py:119-128."""
    torch.manual_seed(3)
    z = torch.randn(*shape)
    want = z.pow(2).sum(-1).pow(power / 2)
    assert (tac.complex_norm(dev(z), power).cpu() - want).abs().max().item() < 1e-5


@pytest.mark.parametrize("new_len", [120, 36, 300])
@pytest.mark.parametrize("shape", [(1, 257, 391), (1, 2, 257, 391)])
def test_apply_filterbank_dense(tac, oc, shape, new_len):
    """tests/test_functional.py:131-141 (shape) + values against the oracle matmul."""
    torch.manual_seed(5)
    spec, fb = torch.randn(*shape), torch.randn(shape[-2], new_len)
    got = tac.apply_filterbank(dev(spec), dev(fb)).cpu()
    want = oc.apply_filterbank(spec, fb)
    assert got.shape == want.shape and got.shape[-2] == new_len and got.shape[-1] == spec.shape[-1]
    assert rel_err(got, want) < REL


# ------------------------------------------------------------------------------------------ stft
def test_cfg1_spectrogram_golden(tac):
    g = golden("cfg1_spectrogram_512_128.npz")
    m = tac.Spectrogram(fft_length=512, hop_length=128).cuda()
    out = m(dev(g["x"])).cpu()
    assert out.shape == (1, 1, 257, 126)
    assert rel_err(out, g["out"]) < REL


def test_stft_reference_config_golden(tac):
    g = golden("stft_512_256.npz")
    out = tac.stft(dev(g["x"]), 512, 256, window=dev(torch.hann_window(512))).cpu()
    assert out.shape == g["out"].shape
    assert rel_err(out, g["out"]) < REL


def test_stft_options_golden(tac):
    g = golden("stft_options.npz")
    x = dev(g["x"])
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
    for tag, kw in cases.items():
        out = tac.STFT(**kw).cuda()(x).cpu()
        assert out.shape == g["out_" + tag].shape, tag
        assert rel_err(out, g["out_" + tag]) < REL, tag


@pytest.mark.parametrize("shape", [(1, 100000), (1, 2, 100000)])
def test_stft_vs_f64(tac, shape):
    """tests/test_functional.py:26-66 with librosa.stft replaced by its float64 restatement."""
    from oracle import f64_chain
    torch.manual_seed(11)
    x = torch.randn(*shape)
    z = tac.stft(dev(x), fft_length=512, hop_length=256, window=dev(torch.hann_window(512))).cpu()
    frames = (x.size(-1) + 2 * 256 - 512 + 256) // 256
    assert z.shape == tuple(x.shape[:-1]) + (257, frames, 2)
    want = f64_chain.stft(x.numpy(), 512, 256)
    got = z.numpy()[..., 0] + 1j * z.numpy()[..., 1]
    assert np.allclose(got, want, atol=1e-4)       # reference asserts 1e-5 on librosa's own fp32 output scale
    assert np.abs(got - want).max() < 5e-5


@pytest.mark.parametrize("fft,hop", [(400, 160), (1200, 300), (441, 110), (6, 2), (1000, 250), (3000, 1000)])
def test_stft_non_power_of_two(tac, oc, fft, hop):
    """functional.py:99 hands any fft_length to torch.stft; sizes that are not a power of two (400 = 25 ms at 16 kHz, odd
    sizes, tiny ones) run the direct-DFT kernel (csrc/stft.cu stft_dft_kernel).  Complex STFT (one- and two-sided),
    spectrogram and the mel chain against the oracle, and their gradients against the oracle under torch autograd."""
    torch.manual_seed(fft)
    x = torch.randn(2, 2, 12000)
    for kw in (dict(), dict(onesided=False, pad_mode="constant"), dict(center=False, normalized=True)):
        got = tac.stft(dev(x), fft, hop, **kw).cpu()
        want = oc.stft(x, fft, hop, **kw)
        assert got.shape == want.shape, (fft, kw)
        assert rel_err(got, want) < REL, (fft, kw, rel_err(got, want))
    got = tac.Spectrogram(fft_length=fft, hop_length=hop, power=1.0).cuda()(dev(x)).cpu()
    assert rel_err(got, oc.spectrogram(x, fft, hop, power=1.0)) < REL
    if fft >= 400:
        m = tac.Sequential(*tac.Melspectrogram(num_mels=40, sample_rate=16000, fft_length=fft, hop_length=hop),
                           tac.AmplitudeToDb()).cuda()
        got = m(dev(x)).cpu()
        want = oc.melspectrogram(x, 40, 16000, to_db=True, fft_length=fft, hop_length=hop)
        assert got.shape == want.shape and (got - want).abs().max().item() < 1e-3
    # gradients: the direct-DFT adjoint (csrc/stft_backward.cu stft_dft_backward_kernel) against the oracle under autograd
    xs = x[:1, :, :3000].contiguous()
    for fn_t, fn_o in ((lambda t: tac.stft(t, fft, hop), lambda t: oc.stft(t, fft, hop)),
                       (lambda t: tac.stft(t, fft, hop, onesided=False, pad_mode="constant"), lambda t: oc.stft(t, fft, hop, onesided=False, pad_mode="constant")),
                       (lambda t: tac.spectrogram(t, fft, hop, power=2.0), lambda t: oc.spectrogram(t, fft, hop, power=2.0)),
                       (lambda t: tac.spectrogram(t, fft, hop, power=1.0, normalized=True), lambda t: oc.spectrogram(t, fft, hop, power=1.0, normalized=True))):
        xo = xs.clone().requires_grad_(True)
        yo = fn_o(xo)
        gy = torch.randn(yo.shape, generator=torch.Generator().manual_seed(fft + 1))
        (want,) = torch.autograd.grad(yo, xo, gy)
        xt = dev(xs).requires_grad_(True)
        (got,) = torch.autograd.grad(fn_t(xt), xt, dev(gy))
        assert rel_err(got.cpu(), want) < REL, (fft, rel_err(got.cpu(), want))


def test_float64_operators(tac, oc):
    """The reference computes in the dtype it is given (functional.py:99 ... :353); double tensors run the double kernels
    of csrc/f64_path.cu.  Every stage against the oracle evaluated in float64 on the CPU, to 1e-10 of the output scale;
    dtype mismatches raise like torch.matmul does in the reference; gradients of double inputs are refused."""
    torch.manual_seed(71)
    x = torch.randn(2, 2, 9000, dtype=torch.float64)
    for fft, hop, kw in ((512, 128, dict()), (400, 160, dict(onesided=False, pad_mode="constant")), (2048, 512, dict(normalized=True)),
                         (256, 64, dict(center=False, win_length=200))):
        got = tac.stft(dev(x), fft, hop, **kw)
        want = oc.stft(x, fft, hop, **kw)
        assert got.dtype == torch.float64 and got.shape == want.shape
        assert (got.cpu() - want).abs().max().item() < 1e-10 * max(1.0, want.abs().max().item()), (fft, kw)
    z = oc.stft(x, 512, 128)
    for power in (1.0, 2.0, 0.7):
        assert torch.allclose(tac.complex_norm(dev(z), power).cpu(), oc.complex_norm(z, power), rtol=1e-12, atol=1e-13)
    mag, ph = tac.magphase(dev(z), 2.0)
    wm, wp = oc.magphase(z, 2.0)
    assert torch.allclose(mag.cpu(), wm, rtol=1e-12) and torch.allclose(ph.cpu(), wp, rtol=1e-12, atol=1e-14)
    assert torch.allclose(tac.angle(dev(z)).cpu(), oc.angle(z), rtol=1e-12, atol=1e-14)
    spec = oc.complex_norm(z, 2.0)
    fb = tac.MelFilterbank(num_freqs=257, num_mels=40, sample_rate=16000).get_filterbank()
    with pytest.raises(RuntimeError):
        tac.apply_filterbank(dev(spec), dev(fb))                # float32 matrix, float64 spectrogram: as torch.matmul
    mel = tac.apply_filterbank(dev(spec), dev(fb.double()))
    want = oc.apply_filterbank(spec, fb.double())
    assert mel.dtype == torch.float64 and torch.allclose(mel.cpu(), want, rtol=1e-12, atol=1e-12)
    db = tac.amplitude_to_db(mel, ref=2.0, amin=1e-6)
    assert torch.allclose(db.cpu(), oc.amplitude_to_db(want, ref=2.0, amin=1e-6), rtol=1e-12, atol=1e-11)
    assert torch.allclose(tac.db_to_amplitude(db, ref=2.0).cpu(), oc.db_to_amplitude(db.cpu(), ref=2.0), rtol=1e-12)
    # modules: .double() moves the buffers, the chain is the reference's
    m = tac.Sequential(*tac.Melspectrogram(num_mels=40, sample_rate=16000, fft_length=512, hop_length=128), tac.AmplitudeToDb()).cuda().double()
    got = m(dev(x)).cpu()
    want = oc.amplitude_to_db(oc.apply_filterbank(oc.complex_norm(oc.stft(x, 512, 128), 2.0), fb.double()))
    assert got.dtype == torch.float64 and (got - want).abs().max().item() < 1e-9
    # mu-law in double
    codes = torch.randint(0, 256, (3, 1000))
    assert torch.equal(tac.mu_law_decoding(dev(codes), 256, torch.float64).cpu(), oc.mu_law_decoding(codes, 256, torch.float64))
    sig = torch.rand(3, 50000, dtype=torch.float64) * 2 - 1
    enc = tac.mu_law_encoding(dev(sig), 256).cpu()
    ref = oc.mu_law_encoding(sig, 256)
    assert (enc != ref).sum().item() <= 1                       # device log1p vs the host's: at most a boundary case
    with pytest.raises(NotImplementedError):
        tac.stft(dev(x).requires_grad_(True), 512, 128)


def test_stft_too_short_raises(tac):
    """tests/test_functional.py:31: reflect padding needs more samples than the pad."""
    with pytest.raises(RuntimeError):
        tac.stft(dev(torch.randn(1, 100)), fft_length=512, hop_length=256)


@pytest.mark.parametrize("fft", [2048])
def test_stft_2048_fast_path(tac, oc, fft):
    torch.manual_seed(13)
    for shape in [(3, 1, 16000), (2, 2, 5000), (1, 4099)]:      # odd length -> gather path everywhere
        x = torch.randn(*shape)
        got = tac.stft(dev(x), fft, 512).cpu()
        want = oc.stft(x, fft, 512)
        assert got.shape == want.shape
        assert rel_err(got, want) < REL
    x = torch.randn(2, 16000)
    xs = dev(torch.randn(2, 16003))[:, 3:]                      # misaligned view -> contiguous copy path
    assert rel_err(tac.stft(xs, fft, 500).cpu(), oc.stft(xs.cpu(), fft, 500)) < REL
    sp = tac.Spectrogram(fft, 512, power=2.0).cuda()(dev(x)).cpu()
    assert rel_err(sp, oc.spectrogram(x, fft, 512, power=2.0)) < REL


@pytest.mark.parametrize("fft,hop", [(2048, 512), (2048, 500), (512, 128), (256, 64), (1024, 256)])
@pytest.mark.parametrize("pad_mode", ["reflect", "replicate", "constant"])
def test_stft_edge_frames_rect_window(tac, oc, fft, hop, pad_mode):
    """Edge frames are bulk-copied and their padding is rebuilt inside shared memory; a rectangular window keeps
    every padded sample visible (a Hann window hides sample 0 of frame 0), short rows make most frames edge frames."""
    torch.manual_seed(31)
    for T in (fft + 4, 3 * fft, 5 * fft + 8):
        x = torch.randn(3, T)
        win = torch.ones(fft)
        got = tac.stft(dev(x), fft, hop, window=dev(win), pad_mode=pad_mode).cpu()
        want = oc.stft(x, fft, hop, window=win, pad_mode=pad_mode)
        assert got.shape == want.shape
        assert rel_err(got, want) < REL, (T, pad_mode)


def test_spectrogram_db_vs_f64(tac):
    """tests/test_layers.py:55-83 with librosa replaced by the float64 restatement (atol 1e-2)."""
    from oracle import f64_chain
    torch.manual_seed(17)
    for shape in [(1, 100000), (1, 2, 100000)]:
        x = torch.randn(*shape)
        model = torch.nn.Sequential(*tac.Spectrogram(512, hop_length=256, window=torch.hann_window(512)),
                                    tac.AmplitudeToDb(ref=1.0, amin=1e-7)).cuda()
        got = model(dev(x)).cpu().numpy()
        want = f64_chain.power_to_db(np.abs(f64_chain.stft(x.numpy(), 512, 256)) ** 2, 1.0, 1e-7)
        assert np.allclose(got, want, atol=1e-2), np.abs(got - want).max()


# ------------------------------------------------------------------------------------------ mel
def test_mel_golden_16k(tac):
    g = golden("mel_16k_2048_512.npz")
    m = tac.Melspectrogram(num_mels=128, sample_rate=16000, fft_length=2048, hop_length=512).cuda()
    out = m(dev(g["x"])).cpu()
    assert out.shape == g["out"].shape == (2, 1, 128, 32)
    assert pure_rel_err(out, g["out"]) < REL
    # unfused, child by child through a plain nn.Sequential: same numbers
    plain = torch.nn.Sequential(*m)
    assert pure_rel_err(plain(dev(g["x"])).cpu(), g["out"]) < REL


def test_meldb_golden_48k(tac):
    g = golden("meldb_48k_2048_512.npz")
    m = tac.Sequential(*tac.Melspectrogram(num_mels=128, sample_rate=48000, fft_length=2048, hop_length=512),
                       tac.AmplitudeToDb()).cuda()
    out = m(dev(g["x"])).cpu()
    assert out.shape == g["out"].shape
    assert (out - g["out"]).abs().max().item() < 1e-3
    plain = torch.nn.Sequential(*m)
    assert (plain(dev(g["x"])).cpu() - g["out"]).abs().max().item() < 1e-3


def test_mel_sweep_golden(tac):
    g = golden("mel_sweep_16k.npz")
    x = dev(g["x"])
    for fft in (256, 512, 1024, 2048, 4096):
        m = tac.Melspectrogram(num_mels=128, sample_rate=16000, fft_length=fft, hop_length=fft // 4).cuda()
        out = m(x).cpu()
        want = g["out_%d" % fft]
        assert out.shape == want.shape
        assert rel_err(out, want) < REL, fft
        nz = want > 0
        assert ((out[nz] - want[nz]).abs() / want[nz]).max().item() < REL, fft
        assert (out[~nz] == 0).all()          # all-zero bands (fft 256 has 13) stay exactly zero


@pytest.mark.parametrize("shape,sr", [((8, 1, 160000), 16000), ((2, 2, 48000), 48000), ((3, 50000), 22050)])
def test_mel_vs_oracle_larger(tac, oc, shape, sr):
    torch.manual_seed(19)
    x = torch.randn(*shape)
    m = tac.Melspectrogram(num_mels=128, sample_rate=sr, fft_length=2048, hop_length=512).cuda()
    got = m(dev(x)).cpu()
    want = oc.melspectrogram(x, 128, sr, fft_length=2048, hop_length=512)
    assert got.shape == want.shape
    assert pure_rel_err(got, want) < REL


def test_mel_nonrandom_signals(tac, oc):
    """uniform noise and sine+noise: near-zero bins and the amin clamp get exercised."""
    torch.manual_seed(23)
    t = torch.arange(64000) / 16000.0
    sig = torch.stack([0.5 * torch.sin(2 * np.pi * 440 * t) + 1e-3 * torch.randn(64000),
                       torch.rand(64000) * 2 - 1, torch.zeros(64000)]).unsqueeze(1)
    m = tac.Sequential(*tac.Melspectrogram(num_mels=128, sample_rate=16000, fft_length=2048, hop_length=512),
                       tac.AmplitudeToDb()).cuda()
    got = m(dev(sig)).cpu()
    want = oc.melspectrogram(sig, 128, 16000, to_db=True, fft_length=2048, hop_length=512)
    # bands 60 dB below the tone are rounding noise of the FFT relative to the peak: there the reference's
    # own fp32 result is only good to a few 1e-3 dB, so judge both against the float64 truth
    from oracle import f64_chain
    fb = oc.mel_filterbank_for(128, 16000, fft_length=2048).numpy()
    truth = f64_chain.power_to_db(f64_chain.melspectrogram(sig.numpy(), fb, 2048, 512) ** 2, 1.0, 1e-7)
    err_ours = np.abs(got.numpy() - truth).max()
    err_ref = np.abs(want.numpy() - truth).max()
    assert err_ours < max(2.0 * err_ref, 1e-3), (err_ours, err_ref)
    assert (got - want).abs().max().item() < 1e-2          # the reference's own dB tolerance (tests/test_layers.py:83)
    loud = want > -20.0
    assert (got[loud] - want[loud]).abs().max().item() < 1e-3
    assert (got[2] == -70.0).all()


def test_full_size_properties_cfg2(tac):
    """BASELINE config 2 at full size: linearity and batch-independence (size-independent checks)."""
    torch.manual_seed(29)
    x = torch.randn(64, 1, 160000, device="cuda")
    m = tac.Melspectrogram(num_mels=128, sample_rate=16000, fft_length=2048, hop_length=512).cuda()
    y = m(x)
    assert y.shape == (64, 1, 128, 313)
    assert torch.isfinite(y).all() and (y >= 0).all()
    y2 = m(2.0 * x)
    assert pure_rel_err(y2.cpu(), (4.0 * y).cpu()) < 1e-5           # power spectrum: scale^2
    ysub = m(x[5:9])
    assert torch.equal(ysub, y[5:9])                                 # a sequence does not see its neighbours
    assert torch.equal(m(x), y)                                      # deterministic


def _log_parity(name, **vals):
    """Measured errors are printed (pytest -s / -rP shows them) and, on the GPU box, appended to
    gpurun_out/parity_errors.jsonl so that a round's evidence run can commit them under profiles/."""
    import json, os
    line = dict(test=name, **{k: (float(v) if isinstance(v, (int, float, np.floating)) else v) for k, v in vals.items()})
    print("PARITY", json.dumps(line))
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "parity_errors.jsonl"), "a") as fh:
            fh.write(json.dumps(line) + "\n")


def _oracle_in_chunks(oc, x, sr, to_db, chunk):
    """The oracle materialises ~60 KB per frame: walk the batch in chunks of sequences."""
    parts = []
    for i in range(0, x.shape[0], chunk):
        parts.append(oc.melspectrogram(x[i:i + chunk], 128, sr, to_db=to_db, fft_length=2048, hop_length=512))
    return torch.cat(parts)


def test_full_size_oracle_cfg2(tac, oc):
    """BASELINE config 2 at FULL size, (64,1,160000): every output value against the oracle (1e-4 relative)."""
    torch.manual_seed(1234)
    x = torch.randn(64, 1, 160000)
    m = tac.Melspectrogram(num_mels=128, sample_rate=16000, fft_length=2048, hop_length=512).cuda()
    got = m(dev(x)).cpu()
    want = _oracle_in_chunks(oc, x, 16000, False, 16)
    assert got.shape == want.shape == (64, 1, 128, 313)
    err = pure_rel_err(got, want)
    _log_parity("cfg2_full (64,1,160000) mel", max_rel=err, values=got.numel())
    assert err < REL


def test_full_size_oracle_cfg3(tac, oc):
    """BASELINE config 3 at FULL size, (256,2,480000) 48 kHz + AmplitudeToDb: every value against the oracle, reference
    walked in chunks of 16 sequences.  Bar: 1e-3 dB absolute (= 2.3e-4 relative in power, i.e. 1.15e-4 in the mel value
    the reference squares, functional.py:291); the measured maximum is logged."""
    torch.manual_seed(1235)
    m = _mel_chain(tac, sr=48000, to_db=True, hop_length=512)
    worst_db, worst_rel, n = 0.0, 0.0, 0
    m_lin = _mel_chain(tac, sr=48000, to_db=False, hop_length=512)
    for start in range(0, 256, 32):                              # GPU in slices too: the CPU side holds one slice
        x = torch.randn(32, 2, 480000)
        xd = dev(x)
        got = m(xd).cpu()
        got_lin = m_lin(xd).cpu()
        want = _oracle_in_chunks(oc, x, 48000, True, 8)
        want_lin = _oracle_in_chunks(oc, x, 48000, False, 8)
        assert got.shape == want.shape == (32, 2, 128, 938)
        worst_db = max(worst_db, (got - want).abs().max().item())
        worst_rel = max(worst_rel, pure_rel_err(got_lin, want_lin))
        n += got.numel()
    _log_parity("cfg3_full (256,2,480000) mel+dB", max_abs_db=worst_db, max_rel_linear=worst_rel, values=n)
    assert worst_db < 1e-3 and worst_rel < REL


def test_full_size_oracle_cfg4_shard(tac, oc):
    """BASELINE config 4: one rank's shard at world 8, (1024,1,160000), in one call (the launch the 8-GPU run makes per
    rank) against the oracle over every value."""
    torch.manual_seed(1236)
    x = torch.randn(1024, 1, 160000)
    m = tac.Melspectrogram(num_mels=128, sample_rate=16000, fft_length=2048, hop_length=512).cuda()
    got = m(dev(x)).cpu()
    want = _oracle_in_chunks(oc, x, 16000, False, 32)
    assert got.shape == want.shape == (1024, 1, 128, 313)
    err = pure_rel_err(got, want)
    _log_parity("cfg4_shard (1024,1,160000) mel", max_rel=err, values=got.numel())
    assert err < REL


def test_full_size_mulaw_cfg5(tac, oc):
    """BASELINE config 5 at FULL size: mu-law codes of (4096,1,240000) uniform[-1,1) equal the oracle's (torch.equal,
    983 M samples, chunked), and decoding them equals the oracle's decode bit for bit."""
    g = torch.Generator().manual_seed(1237)
    bad_enc = bad_dec = 0
    for start in range(0, 4096, 256):
        x = torch.rand(256, 1, 240000, generator=g) * 2 - 1
        codes = tac.mu_law_encoding(dev(x), 256)
        want = oc.mu_law_encoding(x, 256)
        bad_enc += int((codes.cpu() != want).sum())
        dec = tac.mu_law_decoding(codes, 256).cpu()
        bad_dec += int((dec != oc.mu_law_decoding(want, 256)).sum())
    _log_parity("cfg5_full (4096,1,240000) mu-law", encode_mismatches=bad_enc, decode_mismatches=bad_dec, samples=4096 * 240000)
    assert bad_enc == 0 and bad_dec == 0


def test_pair_kernel_matches_single(tac, oc):
    """The two-frames-per-warp kernel (csrc/stft_pair.cu: packed fp32 pairs, tables in tensor memory) against the
    one-frame kernel of round 1 (csrc/stft.cu) and against the oracle: ragged shapes, every padding mode, dB on / off,
    odd frame counts (a last pair with no second frame), hops other than 512, an unaligned view (per-sample gather).
    The two kernels differ only in where the compiler contracts the window multiply into an FMA (a few 1e-7)."""
    lib = tac._cabi.lib()
    torch.manual_seed(61)
    cases = [((3, 1, 16000), 16000, "reflect", False, 512), ((2, 2, 48001), 48000, "reflect", True, 512),
             ((5, 1, 4096), 16000, "constant", False, 512), ((1, 1, 2049), 22050, "replicate", True, 512),
             ((4, 1, 33333), 16000, "circular", False, 512), ((1, 3, 6161), 8000, "reflect", True, 512),
             ((2, 1, 30000), 16000, "reflect", False, 300), ((2, 1, 30000), 16000, "reflect", False, 128),
             ((3, 1, 2048), 16000, "reflect", False, 512)]
    worst = 0.0
    try:
        for shape, sr, pad_mode, db, hop in cases:
            x = torch.randn(*shape)
            m = _mel_chain(tac, sr=sr, to_db=db, hop_length=hop, pad_mode=pad_mode)
            lib.tac_mel_kernel_variant(1)
            single = m(dev(x)).clone()
            lib.tac_mel_kernel_variant(0)
            n0 = lib.tac_launch_count()
            pair = m(dev(x)).clone()
            assert lib.tac_launch_count() - n0 == 1
            want = oc.melspectrogram(x, 128, sr, to_db=db, fft_length=2048, hop_length=hop, pad_mode=pad_mode)
            if db:
                assert (pair - single).abs().max().item() < 1e-4, (shape, pad_mode)
                assert (pair.cpu() - want).abs().max().item() < 1e-3, (shape, pad_mode)
            else:
                err = pure_rel_err(pair.cpu(), single.cpu())
                worst = max(worst, err)
                assert err < 1e-5, (shape, pad_mode, hop)
                assert pure_rel_err(pair.cpu(), want) < REL, (shape, pad_mode, hop)
        xb = torch.randn(3, 1, 20001)[:, :, 1:]
        m = _mel_chain(tac, hop_length=512)
        got = m(dev(xb)[:, :, :]).cpu()                        # storage offset of one float: not 16-byte aligned
        xd = torch.randn(3, 1, 20001, device="cuda")
        xv = xd[:, :, 1:]
        assert pure_rel_err(m(xv).cpu(), oc.melspectrogram(xv.cpu(), 128, 16000, fft_length=2048, hop_length=512)) < REL
        assert got.shape == (3, 1, 128, 40)
    finally:
        lib.tac_mel_kernel_variant(0)
    _log_parity("pair kernel vs one-frame kernel", max_rel=worst)


def test_pair_kernel_tensor_core_pass_matches_oracle(tac, oc):
    """tac_mel_kernel_variant(2): the pair kernel with its second 32-point FFT pass on tcgen05 (3xTF32, operands through
    tensor memory; csrc/stft_pair_tc.cu).  Same cases as above against the oracle and the default kernel; the transform is
    summed in a different order, so agreement is to fp32 rounding (measured 3e-6 relative), not bit for bit.  Chunk lengths
    that are not a multiple of the eight pairs of a round (warps without a pair in the last round) are covered by the
    ragged shapes, full rounds by the (40, 1, 160000) case."""
    lib = tac._cabi.lib()
    torch.manual_seed(67)
    cases = [((3, 1, 16000), 16000, "reflect", False, 512), ((2, 2, 48001), 48000, "reflect", True, 512),
             ((5, 1, 4096), 16000, "constant", False, 512), ((1, 1, 2049), 22050, "replicate", True, 512),
             ((4, 1, 33333), 16000, "circular", False, 512), ((2, 1, 30000), 16000, "reflect", False, 300),
             ((40, 1, 160000), 16000, "reflect", False, 512), ((1, 1, 2048), 16000, "reflect", False, 512)]
    worst = 0.0
    try:
        for shape, sr, pad_mode, db, hop in cases:
            x = torch.randn(*shape)
            m = _mel_chain(tac, sr=sr, to_db=db, hop_length=hop, pad_mode=pad_mode)
            lib.tac_mel_kernel_variant(0)
            pair = m(dev(x)).clone()
            lib.tac_mel_kernel_variant(2)
            n0 = lib.tac_launch_count()
            tc = m(dev(x)).clone()
            assert lib.tac_launch_count() - n0 == 1
            again = m(dev(x))
            assert torch.equal(tc, again), (shape, "not deterministic")
            want = oc.melspectrogram(x, 128, sr, to_db=db, fft_length=2048, hop_length=hop, pad_mode=pad_mode)
            if db:
                assert (tc - pair).abs().max().item() < 2e-4, (shape, pad_mode)
                assert (tc.cpu() - want).abs().max().item() < 1e-3, (shape, pad_mode)
            else:
                err = pure_rel_err(tc.cpu(), pair.cpu())
                worst = max(worst, err)
                assert err < 2e-5, (shape, pad_mode, hop)
                assert pure_rel_err(tc.cpu(), want) < REL, (shape, pad_mode, hop)
    finally:
        lib.tac_mel_kernel_variant(0)
    _log_parity("pair kernel with tcgen05 pass vs default pair kernel", max_rel=worst)


def test_host_pipeline_cfg1(tac):
    g = golden("cfg1_spectrogram_512_128.npz")
    hp = tac.HostPipeline(512, 128, power=1.0)
    out = hp(g["x"])
    assert rel_err(out, g["out"]) < REL
    g2 = golden("meldb_48k_2048_512.npz")
    fb = tac.MelFilterbank(num_freqs=1025, num_mels=128, sample_rate=48000).get_filterbank()
    hp2 = tac.HostPipeline(2048, 512, power=2.0, filterbank=fb, to_db=True)
    assert (hp2(g2["x"]) - g2["out"]).abs().max().item() < 1e-3


def test_melspectrogram_stretch_shapes(tac):
    """tests/test_layers.py:86-106 without TimeStretch (out of scope): hand-composed chain on (4, T)."""
    x = torch.randn(4, 100000, device="cuda")
    fb = tac.MelFilterbank(num_freqs=257, num_mels=128, max_freq=1.0).get_filterbank()
    model = torch.nn.Sequential(tac.STFT(512, hop_length=256), tac.ComplexNorm(power=2.0), tac.ApplyFilterbank(fb)).cuda()
    y = model(x)
    assert y.shape == (4, 128, (100000 + 512 - 512 + 256) // 256)


# ------------------------------------------------------------------------------------------ one-kernel mel path
def _mel_chain(tac, sr=16000, num_mels=128, to_db=False, **stft_kw):
    mods = list(tac.Melspectrogram(num_mels=num_mels, sample_rate=sr, fft_length=2048, **stft_kw))
    if to_db:
        mods.append(tac.AmplitudeToDb())
    return tac.Sequential(*mods).cuda()


def test_fused_mel_is_taken_and_matches_two_kernel_path(tac, oc, monkeypatch):
    """fft 2048 + triangular matrix -> ONE launch (csrc/stft.cu OUT_MEL_FUSED); TAC_MELSPEC_FUSED=0 -> the
    stft + tensor-core filterbank pair.  Both must agree with the oracle."""
    torch.manual_seed(41)
    x = torch.randn(5, 1, 40000)
    m = _mel_chain(tac, hop_length=512)
    lib = tac._cabi.lib()
    m(dev(x))                                                   # plan built, tables uploaded
    n0 = lib.tac_launch_count()
    fused = m(dev(x)).cpu()
    assert lib.tac_launch_count() - n0 == 1
    monkeypatch.setenv("TAC_MELSPEC_FUSED", "0")
    n0 = lib.tac_launch_count()
    pair = m(dev(x)).cpu()
    assert lib.tac_launch_count() - n0 == 2
    want = oc.melspectrogram(x, 128, 16000, fft_length=2048, hop_length=512)
    assert pure_rel_err(fused, want) < REL
    assert pure_rel_err(pair, want) < REL


def test_fused_mel_reference_layout(tac, oc):
    """layout='reference': the memory order behind the reference's `.transpose(-2, -1)` view
    (functional.py:183-184; SURVEY H6 measured strides (..., 1, num_bands))."""
    torch.manual_seed(43)
    x = torch.randn(3, 2, 30000)
    fb = tac.MelFilterbank(num_freqs=1025, num_mels=128, sample_rate=16000).get_filterbank()
    got = tac.functional.melspectrogram(dev(x), dev(fb), 2048, 512)                    # default layout
    want = oc.melspectrogram(x, 128, 16000, fft_length=2048, hop_length=512)
    assert got.shape == want.shape == (3, 2, 128, 59)
    assert got.stride()[-2:] == (1, 128) == want.stride()[-2:]
    assert pure_rel_err(got.cpu(), want) < REL
    same = tac.functional.melspectrogram(dev(x), dev(fb), 2048, 512, layout="contiguous")
    assert same.is_contiguous() and torch.equal(same, got.contiguous())
    module_out = tac.Melspectrogram(num_mels=128, sample_rate=16000, fft_length=2048, hop_length=512).cuda()(dev(x))
    assert module_out.stride() == got.stride() and torch.equal(module_out, got)
    prep = tac.PreparedMelspectrogram(x.shape, "cuda", fb, 2048, 512)
    assert torch.equal(prep(dev(x), prep.empty_output()), got)
    prep_c = tac.PreparedMelspectrogram(x.shape, "cuda", fb, 2048, 512, layout="contiguous")
    assert torch.equal(prep_c(dev(x), prep_c.empty_output()), same)


@pytest.mark.parametrize("power", [1.0, 2.0, 0.7])
def test_fused_mel_other_exponents(tac, oc, power):
    torch.manual_seed(47)
    x = torch.randn(2, 1, 20000)
    fb = tac.MelFilterbank(num_freqs=1025, num_mels=64, sample_rate=22050).get_filterbank()
    got = tac.functional.melspectrogram(dev(x), dev(fb), 2048, 300, power=power).cpu()
    want = oc.apply_filterbank(oc.spectrogram(x, 2048, 300, power=power), fb)
    assert got.shape == want.shape
    assert pure_rel_err(got, want) < REL


@pytest.mark.parametrize("pad_mode", ["reflect", "replicate", "constant", "circular"])
def test_fused_mel_edge_frames_and_options(tac, oc, pad_mode):
    """short rows (most frames touch the padding), rectangular window, normalized, win_length, htk, min_freq"""
    torch.manual_seed(53)
    win = torch.ones(2048)
    fb = tac.functional.create_mel_filter(1025, 96, 200.0, 7000.0, True)
    for T in (2052, 6144, 10248, 4099):                         # 4099: odd length -> per-sample gather path
        x = torch.randn(3, T)
        got = tac.functional.melspectrogram(dev(x), dev(fb), 2048, 512, window=dev(win), pad_mode=pad_mode,
                                            normalized=True, to_db=True, ref=2.0, amin=1e-6).cpu()
        spec = oc.spectrogram(x, 2048, 512, window=win, pad_mode=pad_mode, normalized=True, power=2.0)
        want = oc.amplitude_to_db(oc.apply_filterbank(spec, fb), ref=2.0, amin=1e-6)
        assert got.shape == want.shape
        assert (got - want).abs().max().item() < 1e-3, (T, pad_mode)
    x = torch.randn(2, 9000)
    got = tac.functional.melspectrogram(dev(x), dev(fb), 2048, 512, win_length=1200, center=False).cpu()
    want = oc.apply_filterbank(oc.spectrogram(x, 2048, 512, win_length=1200, center=False, power=2.0), fb)
    assert got.shape == want.shape and pure_rel_err(got, want) < REL


def test_fused_mel_full_size_cfg3_properties(tac):
    """BASELINE config 3 shape per channel pair, reduced batch (16 of 256): dB output finite, deterministic,
    batch-independent, +20*log10(2)*2 dB when the input doubles (the reference squares the power again)."""
    torch.manual_seed(59)
    x = torch.randn(16, 2, 480000, device="cuda")
    m = _mel_chain(tac, sr=48000, to_db=True, hop_length=512)
    y = m(x)
    assert y.shape == (16, 2, 128, 938) and torch.isfinite(y).all()
    assert torch.equal(m(x), y)
    assert torch.equal(m(x[3:5]), y[3:5])
    y2 = m(2.0 * x)
    assert (y2 - y - 40.0 * np.log10(2.0)).abs().max().item() < 1e-3


def test_host_pipeline_mel_fused_and_pair(tac, monkeypatch):
    g = golden("mel_16k_2048_512.npz")
    fb = tac.MelFilterbank(num_freqs=1025, num_mels=128, sample_rate=16000).get_filterbank()
    hp = tac.HostPipeline(2048, 512, power=2.0, filterbank=fb)
    assert pure_rel_err(hp(g["x"]), g["out"]) < REL
    monkeypatch.setenv("TAC_MELSPEC_FUSED", "0")
    hp2 = tac.HostPipeline(2048, 512, power=2.0, filterbank=fb)
    assert pure_rel_err(hp2(g["x"]), g["out"]) < REL


@pytest.mark.parametrize("num_mels,sr", [(40, 16000), (80, 44100), (24, 8000)])
def test_fused_mel_few_wide_bands(tac, oc, num_mels, sr):
    """few, wide triangles: a band collects partial sums of more than four lanes -> the general list walk of
    band_contract instead of its 4 x 4 fast form"""
    torch.manual_seed(61)
    x = torch.randn(2, 1, 30000)
    m = _mel_chain(tac, sr=sr, num_mels=num_mels, hop_length=512)
    got = m(dev(x)).cpu()
    want = oc.melspectrogram(x, num_mels, sr, fft_length=2048, hop_length=512)
    assert got.shape == want.shape and pure_rel_err(got, want) < REL


# ------------------------------------------------------------------------------------------ N4 / N2
def test_pointwise_next_golden(tac):
    g = golden("pointwise_next.npz")
    z = dev(g["z"])
    assert (tac.angle(z).cpu() - g["angle"]).abs().max().item() < 2e-6
    mag, phase = tac.magphase(z, 2.0)
    assert rel_err(mag.cpu(), g["mag_p2"]) < 1e-6 and (phase.cpu() - g["angle"]).abs().max().item() < 2e-6
    assert pure_rel_err(tac.db_to_amplitude(dev(g["db"]), 1.0).cpu(), g["amp_ref1"]) < 2e-6
    assert pure_rel_err(tac.DbToAmplitude(ref=3.0).cuda()(dev(g["db"])).cpu(), g["amp_ref3"]) < 2e-6


def test_db_to_amplitude_known_answers(tac):
    """tests/test_functional.py:144-158, both directions and both round trips."""
    power = torch.tensor([0.000001, 0.0001, 0.1, 1.0, 10.0, 1000000.0])
    db = torch.tensor([-60.0, -40.0, -10.0, 0.0, 10.0, 60.0])
    amp = power.sqrt()
    assert torch.allclose(tac.db_to_amplitude(dev(db), ref=1.0).cpu(), amp, rtol=1e-6, atol=1e-7)
    assert torch.allclose(tac.db_to_amplitude(tac.amplitude_to_db(dev(amp), ref=1.0), ref=1.0).cpu(), amp, rtol=1e-5, atol=1e-7)
    assert torch.allclose(tac.amplitude_to_db(tac.db_to_amplitude(dev(db), ref=1.0), ref=1.0).cpu(), db, atol=1e-5)


def test_magphase_reference_test_formula(tac):
    """tests/test_functional.py:44-47, 62-66: magphase(stft) gives |.| and the angle, and re-assembles."""
    torch.manual_seed(67)
    x = torch.randn(1, 2, 20000)
    z = tac.stft(dev(x), 512, 256, window=dev(torch.hann_window(512)))
    mag, phase = tac.magphase(z)
    back = torch.stack([mag * torch.cos(phase), mag * torch.sin(phase)], dim=-1)
    assert (back - z).abs().max().item() < 1e-4
    assert (mag - tac.complex_norm(z)).abs().max().item() == 0.0
    assert (phase - tac.angle(z)).abs().max().item() == 0.0


@pytest.mark.parametrize("tag,rate", [("0p5", 0.5), ("1p01", 1.01), ("1p3", 1.3), ("2", 2.0)])
def test_phase_vocoder_golden(tac, tag, rate):
    """Against the reference run in float64 (the precision of its own test, tests/test_functional.py:85-93):
    float64 tensors to 1e-8, float32 tensors (float64 inside the kernel) to float32 rounding of the result."""
    g = golden("phase_vocoder.npz")
    spec32, hop = g["spec"], int(g["hop"])
    bins = spec32.shape[-3]
    want = g["out64_" + tag]
    adv = torch.linspace(0, np.pi * hop, bins)[..., None]
    y32 = tac.phase_vocoder(dev(spec32), rate, dev(adv))
    assert y32.dtype == torch.float32 and y32.shape == want.shape
    assert (y32.cpu().double() - want).abs().max().item() < 2e-5       # the reference's own atol vs librosa is 1e-5
    prior = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)                              # as the reference's test does
    try:
        adv64 = torch.linspace(0, np.pi * hop, bins)[..., None]
        y64 = tac.phase_vocoder(dev(spec32.double()), rate, dev(adv64))
    finally:
        torch.set_default_dtype(prior)
    assert y64.dtype == torch.float64 and (y64.cpu() - want).abs().max().item() < 1e-8
    # the reference's float32 run is only good to ~1e-2 after 90 frames; early frames must still agree
    assert (y32.cpu() - g["out32_" + tag])[..., :8, :].abs().max().item() < 5e-3


@pytest.mark.parametrize("shape", [(1, 2, 1025, 400, 2), (1025, 400, 2)])
@pytest.mark.parametrize("rate", [0.5, 1.01, 1.3])
def test_phase_vocoder_reference_test(tac, oc, shape, rate):
    """tests/test_functional.py:69-116 restated: float64, shape ceil(T / rate), values against the librosa
    algorithm (oracle/f64_chain.py) at the reference's atol 1e-5, and against the oracle."""
    from oracle import f64_chain
    torch.manual_seed(71)
    spec = torch.randn(*shape)
    prior = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        spec64 = spec.double()
        adv = torch.linspace(0, np.pi * 256, spec64.shape[-3])[..., None]
        got = tac.phase_vocoder(dev(spec64), rate=rate, phase_advance=dev(adv)).cpu()
        want = oc.phase_vocoder(spec64, rate, adv)
    finally:
        torch.set_default_dtype(prior)
    expected = list(spec.shape)
    expected[-2] = int(np.ceil(expected[-2] / rate))
    assert list(got.shape) == expected and got.dim() == spec.dim()
    assert (got - want).abs().max().item() < 1e-7
    mono = spec64[(0,) * (spec.dim() - 3)].numpy()
    lib = f64_chain.phase_vocoder(mono[..., 0] + 1j * mono[..., 1], rate, 256)
    g0 = got[(0,) * (spec.dim() - 3)].numpy()
    assert np.allclose(g0[..., 0] + 1j * g0[..., 1], lib, atol=1e-5)


@pytest.mark.parametrize("shape", [(1, 2, 100000), (4, 100000)])
def test_melspectrogram_stretch_pipeline(tac, oc, shape):
    """tests/test_layers.py:86-106 as written: STFT -> TimeStretch(0.7) -> ComplexNorm(2) -> ApplyFilterbank;
    shape as the reference asserts, values against the oracle chain evaluated in float64 for the vocoder."""
    torch.manual_seed(73)
    x = torch.randn(*shape)
    fft, hop, rate = 512, 256, 0.7
    fb = tac.MelFilterbank(num_freqs=fft // 2 + 1, num_mels=128, max_freq=1.0).get_filterbank()
    model = torch.nn.Sequential(tac.STFT(fft, hop_length=hop), tac.TimeStretch(hop_length=hop, num_freqs=fft // 2 + 1, fixed_rate=rate),
                                tac.ComplexNorm(power=2.0), tac.ApplyFilterbank(fb)).cuda()
    y = model(dev(x)).cpu()
    frames = (x.size(-1) + 2 * (fft // 2) - fft + hop) // hop
    assert y.size(-2) == 128 and y.size(-1) == int(np.ceil(frames / rate))
    z = oc.stft(x, fft, hop)
    adv = torch.linspace(0, np.pi * hop, fft // 2 + 1)[..., None]
    steps = torch.arange(0, z.size(-2), rate)                       # float32 steps, as the float32 pipeline has them
    stretched = oc.phase_vocoder(z.double(), rate, adv.double())
    want = oc.apply_filterbank(oc.complex_norm(stretched.float(), 2.0), fb)
    assert stretched.size(-2) == steps.numel()
    assert rel_err(y, want) < 5e-4                                   # |.|^2 does not see the phase; magnitudes interpolate in fp32


# ------------------------------------------------------------------------------------------ empty / ragged / strided inputs
def test_empty_batch_everywhere(tac):
    """Zero sequences: every entry point returns an empty tensor of the reference's shape without launching."""
    x = torch.zeros(0, 1, 4000, device="cuda")
    assert tac.stft(x, 512, 128).shape == (0, 1, 257, 32, 2)
    assert tac.Spectrogram(512, 128).cuda()(x).shape == (0, 1, 257, 32)
    assert tac.Melspectrogram(num_mels=128, sample_rate=16000, fft_length=2048, hop_length=512).cuda()(x).shape == (0, 1, 128, 8)
    assert tac.Melspectrogram(num_mels=40, sample_rate=16000, fft_length=512, hop_length=128).cuda()(x).shape == (0, 1, 40, 32)
    assert tac.mu_law_encoding(torch.zeros(0, 7, device="cuda")).shape == (0, 7)
    assert tac.mu_law_decoding(torch.zeros(0, 7, dtype=torch.int64, device="cuda")).shape == (0, 7)
    assert tac.amplitude_to_db(torch.zeros(0, 3, device="cuda")).shape == (0, 3)


@pytest.mark.parametrize("n_samples", [2049, 16001, 40003, 3 * 512 + 2048])
def test_fused_mel_ragged_lengths(tac, oc, n_samples):
    """Lengths that are not a multiple of 4 / of the hop: rows lose their 16-byte alignment, so interior frames take
    the gather instead of the bulk copy; the last frame ends exactly at / before the end of the row."""
    torch.manual_seed(n_samples)
    x = torch.randn(3, 1, n_samples)
    m = _mel_chain(tac, hop_length=512)
    want = oc.melspectrogram(x, 128, 16000, fft_length=2048, hop_length=512)
    got = m(dev(x)).cpu()
    assert got.shape == want.shape
    assert pure_rel_err(got, want) < REL


def test_fused_mel_tiny_and_strided_inputs(tac, oc):
    """Fewer frames than warps / than SMs (most CTAs get an empty chunk), 1-D and 2-D inputs (functional.py:89-91
    flattens whatever leads), and a batch that is a strided view (every other row of a larger tensor)."""
    torch.manual_seed(77)
    m = _mel_chain(tac, hop_length=512)
    for shape in [(2048,), (1, 2500), (2, 1, 5000), (1, 1, 1, 9000)]:
        x = torch.randn(*shape)
        want = oc.melspectrogram(x, 128, 16000, fft_length=2048, hop_length=512)
        got = m(dev(x)).cpu()
        assert got.shape == want.shape, shape
        assert pure_rel_err(got, want) < REL, shape
    big = torch.randn(8, 1, 30000)
    view = dev(big)[::2]                                    # stride 2 * 30000 between sequences
    assert not view.is_contiguous()
    want = oc.melspectrogram(big[::2], 128, 16000, fft_length=2048, hop_length=512)
    assert pure_rel_err(m(view).cpu(), want) < REL
    st = tac.stft(view, 2048, 512).cpu()
    assert rel_err(st, oc.stft(big[::2], 2048, 512)) < REL


def test_repeated_calls_are_bit_identical(tac):
    """Fixed summation order everywhere on the forward path: the same input gives the same bits, call after call and
    through every entry point (module, functional, prepared)."""
    torch.manual_seed(3)
    x = dev(torch.randn(6, 2, 50000))
    m = _mel_chain(tac, to_db=True, hop_length=512)
    a = m(x)
    for _ in range(3):
        assert torch.equal(m(x), a)


def test_cuda_graph_capture_and_replay(tac, oc):
    """`PreparedMelspectrogram` is a single C-ABI call into caller-owned buffers with no host synchronisation, so it can be
    captured into a CUDA graph and replayed: one-kernel path (fft 2048), the small-size one-kernel path (fft 512) and the
    mu-law pair; replays on new input contents are bit-identical to the eager calls."""
    torch.manual_seed(83)
    for fft, hop, shape in ((2048, 512, (4, 1, 16000)), (512, 128, (2, 2, 8000))):
        fb = tac.MelFilterbank(num_freqs=fft // 2 + 1, num_mels=64, sample_rate=16000).get_filterbank()
        prep = tac.PreparedMelspectrogram(shape, "cuda", fb, fft, hop, to_db=True)
        x = torch.randn(*shape, device="cuda")
        out = prep.empty_output()
        prep(x, out)                                            # warm-up outside the capture (function attributes, plan upload)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for _ in range(3):                                  # several launches in one graph
                y = prep(x, out)
        for _ in range(2):
            x.copy_(torch.randn(*shape, device="cuda"))
            graph.replay()
            torch.cuda.synchronize()
            got = y.clone()
            eager = prep(x, prep.empty_output())
            assert torch.equal(got, eager)
        want = oc.melspectrogram(x.cpu(), 64, 16000, to_db=True, fft_length=fft, hop_length=hop)
        assert (got.cpu() - want).abs().max().item() < 1e-3


def test_hpss_golden(tac, oc):
    """beta_hpss.py:37-129 (unexported beta module of the reference): medians are selections, so harmonic / percussive
    parts and masks are bit-exact for powers 1 and 2 (soft and hard masks, mask_only, several kernel sizes); a general
    power goes through the device powf (1e-6); a NaN in a window gives NaN like torch.median; module repr and the
    reference's failure modes (kernel halves that differ, padding not smaller than the axis)."""
    from torchaudio_contrib_b200.beta_hpss import HPSS, hpss
    g = golden("hpss.npz")
    x = dev(g["x"])
    cases = {"k31_p2": dict(kernel_size=31, power=2.0), "k17_p1": dict(kernel_size=17, power=1.0),
             "k31_hard": dict(kernel_size=31, power=2.0, hard=True), "k9_maskonly": dict(kernel_size=(9, 9), power=2.0, mask_only=True)}
    for tag, kw in cases.items():
        out = hpss(x, **kw)
        assert len(out) == 4
        for i, o in enumerate(out):
            if o is None:
                assert "%s_%d" % (tag, i) not in g
                continue
            want = g["%s_%d" % (tag, i)]
            assert o.dtype == want.dtype and torch.equal(o.cpu(), want), (tag, i)
    out = HPSS(kernel_size=5, power=0.7).cuda()(x)
    for i, o in enumerate(out):
        assert torch.allclose(o.cpu(), g["k5_p07_%d" % i], rtol=2e-6, atol=1e-7), i
    out = hpss(dev(g["xn"]), 5, 1.0)
    for i, o in enumerate(out):
        want = g["nan_%d" % i]
        assert torch.equal(torch.isnan(o.cpu()), torch.isnan(want)) and torch.equal(torch.nan_to_num(o.cpu()), torch.nan_to_num(want)), i
    assert repr(HPSS(7, 1.0, True, False)) == "HPSS(kernel_size=7, power=1.0, hard=True, mask_only=False)"
    # a larger, non-square case against the oracle, kernel 63 (the widest window built)
    torch.manual_seed(97)
    big = torch.rand(3, 1, 257, 130) * 5.0
    for kw in (dict(kernel_size=63, power=2.0), dict(kernel_size=3, power=1.0, hard=True)):
        got, want = hpss(dev(big), **kw), oc.hpss(big, **kw)
        for a, b in zip(got, want):
            assert torch.equal(a.cpu(), b)
    with pytest.raises(RuntimeError):
        hpss(x, (31, 5))                                           # the reference's slices only fit when both halves agree
    with pytest.raises(RuntimeError):
        hpss(dev(torch.rand(1, 1, 10, 50)), 31)                    # reflect padding 15 >= 10 bins
    with pytest.raises(TypeError):
        hpss(x, 31.0)


def test_module_chain_prepared_call_follows_its_buffers(tac, oc):
    """The module chain keeps a prepared call per (shape, device, window, matrix, options) (layers.FusedSequential._prepared_mel);
    it must notice when what it was prepared from changes: an in-place edit of the filterbank, a new window, another shape,
    another device-side dtype path (float64 input), the dB module's parameters."""
    torch.manual_seed(89)
    x = torch.randn(3, 1, 20000)
    mel = tac.Melspectrogram(num_mels=64, sample_rate=16000, fft_length=2048, hop_length=512).cuda()
    y0 = mel(dev(x))
    assert pure_rel_err(y0.cpu(), oc.melspectrogram(x, 64, 16000, fft_length=2048, hop_length=512)) < REL
    assert torch.equal(mel(dev(x)), y0)                                   # steady state: the cached call
    mel[2].filterbank.mul_(2.0)                                           # version counter bumps
    assert pure_rel_err(mel(dev(x)).cpu(), 2.0 * y0.cpu()) < 1e-6
    mel[2].filterbank.mul_(0.5)
    w = torch.hann_window(2048) ** 2
    mel[0].window = dev(w)                                                # another window buffer
    want = oc.melspectrogram(x, 64, 16000, fft_length=2048, hop_length=512, window=w)
    assert pure_rel_err(mel(dev(x)).cpu(), want) < REL
    x2 = torch.randn(2, 2, 9000)                                          # another shape
    assert pure_rel_err(mel(dev(x2)).cpu(), oc.melspectrogram(x2, 64, 16000, fft_length=2048, hop_length=512, window=w)) < REL
    chain = tac.Sequential(*mel, tac.AmplitudeToDb(ref=2.0, amin=1e-6)).cuda()
    got = chain(dev(x2)).cpu()
    assert (got - oc.melspectrogram(x2, 64, 16000, to_db=True, ref=2.0, amin=1e-6, fft_length=2048, hop_length=512, window=w)).abs().max().item() < 1e-3
    chain[3].ref = 1.0                                                     # the dB parameters are part of the key
    got = chain(dev(x2)).cpu()
    assert (got - oc.melspectrogram(x2, 64, 16000, to_db=True, ref=1.0, amin=1e-6, fft_length=2048, hop_length=512, window=w)).abs().max().item() < 1e-3
    xg = dev(x2).requires_grad_(True)                                      # autograd bypasses the prepared call
    assert mel(xg).requires_grad


def test_host_pipeline_other_sizes(tac, oc):
    """The host-buffer entry (tac_pipeline_*) at an fft length that is not a power of two (direct-DFT kernel -> tensor-core
    filterbank, two kernels per slice) and at 1024 (one-kernel range-plan path is not the pipeline's: it runs K1 -> K2 too)."""
    torch.manual_seed(101)
    x = torch.randn(5, 1, 20000)
    for fft, hop, mels in ((400, 160, 40), (1024, 256, 64)):
        fb = tac.MelFilterbank(num_freqs=fft // 2 + 1, num_mels=mels, sample_rate=16000).get_filterbank()
        hp = tac.HostPipeline(fft, hop, power=2.0, filterbank=fb, to_db=True)
        got = hp(x)
        want = oc.melspectrogram(x, mels, 16000, to_db=True, fft_length=fft, hop_length=hop)
        assert got.shape == want.shape and (got - want).abs().max().item() < 1e-3, fft
        sp = tac.HostPipeline(fft, hop, power=1.0)
        assert rel_err(sp(x), oc.spectrogram(x, fft, hop, power=1.0)) < REL, fft


def test_c_abi_without_python(tac, oc, tmp_path):
    """examples/c_abi_mulaw.c: a C program that links libtac_b200.so, takes the mu-law decision levels from
    tac_mulaw_tables_host (shipped in the library) and encodes / decodes a file of samples -- no Python, no torch in that
    process.  Its codes and decoded values equal the reference chain's bit for bit (functional.py:317-354)."""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None or not os.path.exists("/usr/local/cuda/include/cuda_runtime.h"):
        pytest.skip("no C compiler / CUDA headers on this box")
    exe = str(tmp_path / "c_abi_mulaw")
    libdir = os.path.join(root, "torchaudio_contrib_b200", "lib")
    subprocess.run([cc, "-O2", "-I", os.path.join(root, "include"), "-I", "/usr/local/cuda/include",
                    os.path.join(root, "examples", "c_abi_mulaw.c"), "-o", exe, "-L", libdir, "-ltac_b200",
                    "-L", "/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath," + libdir], check=True)
    torch.manual_seed(103)
    x = torch.cat([torch.rand(200000) * 2 - 1, 2 * (torch.randn(50000) - 0.5), torch.tensor([0.0, -0.0, 1.0, -1.0, 1e-30, 3e38])])
    x.numpy().tofile(str(tmp_path / "x.f32"))
    res = subprocess.run([exe, str(tmp_path / "x.f32"), str(tmp_path / "codes.i64"), str(tmp_path / "dec.f32")],
                         check=True, capture_output=True, text=True)
    assert "through the C ABI" in res.stdout
    codes = torch.from_numpy(np.fromfile(str(tmp_path / "codes.i64"), dtype=np.int64))
    dec = torch.from_numpy(np.fromfile(str(tmp_path / "dec.f32"), dtype=np.float32))
    want = oc.mu_law_encoding(x, 256)
    assert torch.equal(codes, want)
    inside = (want >= 0) & (want < 256)
    assert torch.equal(dec[inside], oc.mu_law_decoding(want, 256)[inside])
