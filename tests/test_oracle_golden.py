"""The oracle against (a) the committed vectors generated from the unmodified reference
(oracle/gen_golden.py), (b) the reference tests' own known-answer vectors and formulas, and
(c) the float64 restatement that stands in for librosa.  CPU only."""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import f64_chain, ref_chain as oc


def close(a, b, rtol=2e-6, atol=1e-6):
    return torch.allclose(a, b, rtol=rtol, atol=atol)


def test_cfg1_and_stft_fixtures():
    g = golden("cfg1_spectrogram_512_128.npz")
    assert close(oc.spectrogram(g["x"], 512, 128), g["out"], atol=1e-5)
    g = golden("stft_512_256.npz")
    assert close(oc.stft(g["x"], 512, 256, window=torch.hann_window(512)), g["out"], atol=1e-5)


def test_stft_option_fixtures():
    g = golden("stft_options.npz")
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
        out = oc.stft(g["x"], **kw)
        assert out.shape == g["out_" + tag].shape, tag
        assert close(out, g["out_" + tag], atol=1e-5), tag
        # the float64 restatement agrees with the reference's fp32 torch path
        kw64 = dict(kw)
        n_fft = kw64.pop("fft_length")
        z = f64_chain.stft(g["x"].numpy(), n_fft, kw64.pop("hop_length", None), **kw64)
        ref = g["out_" + tag].numpy()
        assert np.abs(z - (ref[..., 0] + 1j * ref[..., 1])).max() < 2e-4, tag


def test_mel_fixtures():
    g = golden("mel_16k_2048_512.npz")
    assert close(oc.melspectrogram(g["x"], 128, 16000, fft_length=2048, hop_length=512), g["out"], rtol=1e-5, atol=1e-3)
    g = golden("meldb_48k_2048_512.npz")
    assert close(oc.melspectrogram(g["x"], 128, 48000, to_db=True, fft_length=2048, hop_length=512), g["out"], atol=1e-4)
    g = golden("mel_sweep_16k.npz")
    for fft in (256, 512, 1024, 2048, 4096):
        out = oc.melspectrogram(g["x"], 128, 16000, fft_length=fft, hop_length=fft // 4)
        assert close(out, g["out_%d" % fft], rtol=1e-5, atol=1e-3), fft


def test_filterbank_fixtures_bit_exact():
    g = golden("filterbanks.npz")
    assert torch.equal(oc.mel_filterbank_for(128, 16000, fft_length=2048), g["fb_16k_1025x128"])
    assert torch.equal(oc.mel_filterbank_for(128, 48000, fft_length=2048), g["fb_48k_1025x128"])
    for fft in (256, 512, 1024, 4096):
        assert torch.equal(oc.mel_filterbank_for(128, 16000, fft_length=fft), g["fb_16k_%dx128" % (fft // 2 + 1)])
    assert torch.equal(oc.create_mel_filter(1025, 40, 30.0, 22050 // 2, True), g["fb_htk_22k_1025x40"])
    assert torch.equal(oc.create_mel_filter(257, 128, 0.0, 1.0, False), g["fb_maxfreq1_257x128"])
    fb = g["fb_16k_1025x128"]
    assert ((fb != 0).sum(1) <= 2).all()                       # a bin feeds at most two bands
    assert int(((g["fb_16k_129x128"] != 0).sum(0) == 0).sum()) == 13   # fft 256: 13 empty bands


def test_stage_fixtures():
    g = golden("stages.npz")
    assert close(oc.complex_norm(g["z"], 0.7), g["norm_p07"])
    assert close(oc.complex_norm(g["z"], 1.0), g["norm_p1"])
    assert close(oc.apply_filterbank(g["norm_p1"], g["fb"]), g["filtered"], rtol=1e-5, atol=1e-4)
    assert close(oc.amplitude_to_db(g["norm_p1"], 2.0, 1e-5), g["db"], atol=1e-5)


def test_mulaw_fixtures_bit_exact():
    g = golden("mulaw.npz")
    assert torch.equal(oc.mu_law_encoding(g["x"], 256), g["enc256"])
    assert torch.equal(oc.mu_law_encoding(g["x"], 64), g["enc64"])
    assert torch.equal(oc.mu_law_decoding(torch.arange(256), 256), g["dec256"])
    assert torch.equal(oc.mu_law_decoding(torch.arange(64), 64), g["dec64"])
    assert torch.equal(oc.mu_law_encoding(g["dec256"], 256), g["roundtrip256"])
    assert torch.equal(g["roundtrip256"], torch.arange(256))      # encode(decode(i)) == i (tests/test_functional.py:195-199)


# ---- the reference's own known-answer tests, restated against the oracle -----------------------
def test_amplitude_db_known_answers():
    """tests/test_functional.py:144-158."""
    power = torch.tensor([0.000001, 0.0001, 0.1, 1.0, 10.0, 1000000.0])
    db = torch.tensor([-60.0, -40.0, -10.0, 0.0, 10.0, 60.0])
    assert (oc.amplitude_to_db(power.sqrt(), ref=1.0) - db).abs().max() < 1e-5


@pytest.mark.parametrize("power", [1, 2, 0.7])
def test_complex_norm_formula(power):
    """tests/test_functional.py:119-128."""
    z = torch.randn(3, 1025, 40, 2, generator=torch.Generator().manual_seed(1))
    assert (oc.complex_norm(z, power) - z.pow(2).sum(-1).pow(power / 2)).abs().max() < 1e-5


@pytest.mark.parametrize("shape", [(1, 100000), (1, 2, 100000)])
def test_stft_against_float64(shape):
    """tests/test_functional.py:26-66 with librosa.stft -> its float64 restatement, atol 1e-5 kept."""
    x = torch.randn(*shape, generator=torch.Generator().manual_seed(2))
    z = oc.stft(x, 512, 256, window=torch.hann_window(512))
    frames = (x.size(-1) + 2 * 256 - 512 + 256) // 256
    assert z.shape == tuple(x.shape[:-1]) + (257, frames, 2)
    want = f64_chain.stft(x.numpy(), 512, 256)
    got = z.numpy()[..., 0] + 1j * z.numpy()[..., 1]
    assert np.allclose(got, want, atol=1e-4)
    assert f64_chain.num_frames(x.size(-1), 512, 256) == frames


def test_stft_short_input_raises():
    """tests/test_functional.py:31 (strict xfail raises=RuntimeError)."""
    with pytest.raises(RuntimeError):
        oc.stft(torch.randn(1, 100), 512, 256)


def test_spectrogram_db_against_float64():
    """tests/test_layers.py:55-83, atol 1e-2."""
    x = torch.randn(1, 2, 50000, generator=torch.Generator().manual_seed(3))
    got = oc.amplitude_to_db(oc.spectrogram(x, 512, 256, window=torch.hann_window(512)), 1.0, 1e-7).numpy()
    want = f64_chain.power_to_db(np.abs(f64_chain.stft(x.numpy(), 512, 256)) ** 2, 1.0, 1e-7)
    assert np.allclose(got, want, atol=1e-2)


# ---- rows SURVEY 8(f) marks "next": N2 phase vocoder, N4 db_to_amplitude / angle / magphase --------------
def test_pointwise_next_fixtures_bit_exact():
    g = golden("pointwise_next.npz")
    assert torch.equal(oc.angle(g["z"]), g["angle"])
    mag, phase = oc.magphase(g["z"], 2.0)
    assert torch.equal(mag, g["mag_p2"]) and torch.equal(phase, g["angle"])
    assert torch.equal(oc.db_to_amplitude(g["db"], 1.0), g["amp_ref1"])
    assert torch.equal(oc.db_to_amplitude(g["db"], 3.0), g["amp_ref3"])


def test_db_to_amplitude_known_answers_and_round_trips():
    """tests/test_functional.py:144-158: dB [-60..60] -> amplitude sqrt(power), and both round trips."""
    power = torch.tensor([0.000001, 0.0001, 0.1, 1.0, 10.0, 1000000.0])
    db = torch.tensor([-60.0, -40.0, -10.0, 0.0, 10.0, 60.0])
    amp = power.sqrt()
    assert torch.allclose(oc.db_to_amplitude(db, ref=1.0), amp, rtol=1e-6, atol=1e-7)
    assert torch.allclose(oc.db_to_amplitude(oc.amplitude_to_db(amp, ref=1.0), ref=1.0), amp, rtol=1e-5, atol=1e-7)
    assert torch.allclose(oc.amplitude_to_db(oc.db_to_amplitude(db, ref=1.0), ref=1.0), db, atol=1e-5)


@pytest.mark.parametrize("tag,rate", [("0p5", 0.5), ("1p01", 1.01), ("1p3", 1.3), ("2", 2.0)])
def test_phase_vocoder_fixtures(tag, rate):
    """The oracle reproduces the reference's float64 and float32 runs bit for bit, and the float64 run agrees
    with the librosa algorithm restated in numpy (the oracle of tests/test_functional.py:100-116, atol 1e-5)."""
    g = golden("phase_vocoder.npz")
    spec32, hop = g["spec"], int(g["hop"])
    bins = spec32.shape[-3]
    prior = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        adv = torch.linspace(0, np.pi * hop, bins)[..., None]
        y64 = oc.phase_vocoder(spec32.double(), rate, adv)
    finally:
        torch.set_default_dtype(prior)
    assert torch.equal(y64, g["out64_" + tag])
    adv32 = torch.linspace(0, np.pi * hop, bins)[..., None]
    assert torch.equal(oc.phase_vocoder(spec32, rate, adv32), g["out32_" + tag])
    want_frames = int(np.ceil(spec32.shape[-2] / rate))
    assert y64.shape == spec32.shape[:-2] + (want_frames, 2)                      # tests/test_functional.py:95-99
    z = spec32[0, 1].double().numpy()
    lib = f64_chain.phase_vocoder(z[..., 0] + 1j * z[..., 1], rate, hop)
    got = y64[0, 1].numpy()
    assert np.allclose(got[..., 0] + 1j * got[..., 1], lib, atol=1e-7)


def test_hpss_fixtures_bit_exact():
    """oracle.ref_chain.hpss against tests/golden/hpss.npz (oracle/gen_golden.py hpss: the unmodified beta_hpss.py)."""
    g = golden("hpss.npz")
    cases = {"k31_p2": dict(kernel_size=31, power=2.0), "k17_p1": dict(kernel_size=17, power=1.0),
             "k5_p07": dict(kernel_size=5, power=0.7), "k31_hard": dict(kernel_size=31, power=2.0, hard=True),
             "k9_maskonly": dict(kernel_size=(9, 9), power=2.0, mask_only=True)}
    for tag, kw in cases.items():
        out = oc.hpss(g["x"], **kw)
        for i, o in enumerate(out):
            if o is None:
                assert "%s_%d" % (tag, i) not in g
            else:
                assert torch.equal(o, g["%s_%d" % (tag, i)]), (tag, i)
