"""CPU checks of the mathematics behind the adjoint kernels (csrc/stft.cu `stft2048_backward_kernel`,
csrc/stft_backward.cu `overlap_add_kernel`): numpy float64 models that follow the kernels' formulas, against torch
autograd through the reference's operators (torch.stft's framing / padding, rfft, |.|^p).  No GPU, no library call."""
import numpy as np
import pytest
import torch


@pytest.mark.parametrize("power", [2.0, 1.0, 0.7])
def test_collapsed_untangle_gain_retangle(power):
    """Zt_k = 2[(sigma + delta Im W^k) Zh_k + i delta Re W^k conj Zh_{C-k}] (k = 0 doubled), then an inverse complex FFT of
    size C = N/2 whose (Re, Im) pairs are the (even, odd) samples: equals d/dx sum_k g_k |rfft(x w)_k|^p."""
    rng = np.random.default_rng(int(10 * power))
    N, C = 2048, 1024
    x, g = rng.standard_normal(N), rng.standard_normal(C + 1)
    w = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(N) / N)

    xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    X = torch.fft.rfft(xt * torch.tensor(w))
    mag2 = X.real ** 2 + X.imag ** 2
    (torch.tensor(g) * (mag2 if power == 2.0 else mag2 ** (power / 2))).sum().backward()
    want = xt.grad.numpy()

    xw = 0.5 * x * w                                         # the kernels' window table carries the 1/2
    zh = np.fft.fft(xw[0::2] + 1j * xw[1::2])                # Zh: what both register passes of the forward kernel produce
    k = np.arange(C)
    wk = np.exp(-2j * np.pi * k / N)
    q = zh[(C - k) % C]
    p_, q_ = zh, np.conj(q)
    s_, d_ = p_ + q_, -1j * wk * (p_ - q_)                   # X_k = S + D, conj X_{C-k} = S - D

    def gain(n2):                                            # (p / 2) |X|^(p - 2)
        return np.ones_like(n2) if power == 2.0 else 0.5 * power * n2 ** (0.5 * power - 1.0)

    hk = g[k] * gain(np.abs(s_ + d_) ** 2)
    hm = g[C - k] * gain(np.abs(s_ - d_) ** 2)
    sigma, delta = hk + hm, hk - hm
    zt = 2 * ((sigma + delta * wk.imag) * p_ + 1j * delta * wk.real * q_)
    zt[0] *= 2                                               # H_0 = G_0, not G_0 / 2
    y = np.fft.ifft(zt) * C                                  # unnormalised inverse
    grad = np.empty(N)
    grad[0::2], grad[1::2] = y.real, y.imag
    grad *= w
    assert np.abs(grad - want).max() <= 1e-12 * np.abs(want).max()


def _ola_model(frames_ws, n_samples, n_fft, hop, pad, pad_mode):
    """overlap_add_kernel (+ overlap_add_edges_kernel) statement for statement."""
    n_frames = frames_ws.shape[0]

    def padded(qpos):
        t_lo = qpos - n_fft + 1
        t_lo = 0 if t_lo <= 0 else (t_lo + hop - 1) // hop
        t_hi = min(qpos // hop, n_frames - 1)
        return sum(frames_ws[t, qpos - t * hop] for t in range(t_lo, t_hi + 1))

    out = np.zeros(n_samples)
    for j in range(n_samples):
        acc = padded(j + pad)
        if pad > 0:
            if pad_mode == "reflect":
                if 1 <= j <= pad:
                    acc += padded(pad - j)
                if n_samples - 1 - pad <= j <= n_samples - 2:
                    acc += padded(pad + 2 * (n_samples - 1) - j)
            elif pad_mode == "circular":
                if j >= n_samples - pad:
                    acc += padded(j + pad - n_samples)
                if j < pad:
                    acc += padded(j + pad + n_samples)
        out[j] = acc
    if pad > 0 and pad_mode == "replicate":
        out[0] += sum(padded(qpos) for qpos in range(pad))
        out[n_samples - 1] += sum(padded(pad + n_samples + qpos) for qpos in range(pad))
    return out


@pytest.mark.parametrize("pad_mode", ["reflect", "constant", "replicate", "circular"])
@pytest.mark.parametrize("n_samples,n_fft,hop,center", [(100, 32, 8, True), (77, 32, 12, True), (64, 16, 16, True), (90, 32, 8, False)])
def test_overlap_add_is_the_adjoint_of_padding_and_framing(pad_mode, n_samples, n_fft, hop, center):
    rng = np.random.default_rng(n_samples + hop)
    pad = n_fft // 2 if center else 0
    x = torch.tensor(rng.standard_normal(n_samples), dtype=torch.float64, requires_grad=True)
    xp = torch.nn.functional.pad(x.view(1, 1, -1), (pad, pad), mode=pad_mode).view(-1) if pad else x
    fr = xp.unfold(0, n_fft, hop)                            # torch.stft's framing of the padded signal
    gfr = torch.tensor(rng.standard_normal(tuple(fr.shape)))
    (fr * gfr).sum().backward()
    got = _ola_model(gfr.numpy(), n_samples, n_fft, hop, pad, pad_mode)
    assert np.abs(got - x.grad.numpy()).max() < 1e-12


def test_in_place_pairing_of_the_backward_kernel():
    """stft2048_backward_kernel's combine loop (csrc/stft.cu): step (k1, 31 - k1) overwrites registers k1 and 31 - k1 of
    every lane; partners come by shuffle from lane 32 - l (registers 31 - k1 and k1), except lane 0, which pairs with its
    own registers 32 - k1 (carried from the previous step, already overwritten) and k1 + 1.  Model: registers hold the
    bin index they were loaded with, or -1 once overwritten; every bin k must meet bin (1024 - k) mod 1024, untouched."""
    reg = [[32 * k1 + lane for k1 in range(32)] for lane in range(32)]        # reg[lane][k1]
    met = {}
    carry = reg[0][0]
    for k1 in range(16):
        za = [reg[lane][k1] for lane in range(32)]
        zb = [reg[lane][31 - k1] for lane in range(32)]
        qa = [zb[(32 - lane) & 31] for lane in range(32)]                     # shuffles read before anything is written
        qb = [za[(32 - lane) & 31] for lane in range(32)]
        qa[0], qb[0] = carry, reg[0][k1 + 1]
        carry = zb[0]
        for lane in range(32):
            met[za[lane]] = qa[lane]
            met[zb[lane]] = qb[lane]
            reg[lane][k1] = reg[lane][31 - k1] = -1
    assert sorted(met) == list(range(1024))
    for k, partner in met.items():
        assert partner == (1024 - k) % 1024, (k, partner)


def test_phase_vocoder_gather_ranges_partition_the_output_steps():
    """The backward kernel of the phase vocoder gathers, per input frame i, the output steps j with idx0[j] == i and those
    with idx1[j] == i; both index tables are monotone, so the steps form contiguous ranges (functional._phase_vocoder_ranges).
    Checked here on the CPU for stretch and compress rates: every step lies in exactly the range of its frame."""
    import torchaudio_contrib_b200.functional as F
    for n_in, rate in ((41, 0.7), (50, 1.3), (37, 2.0), (12, 0.3), (100, 1.01), (5, 3.7)):
        steps = torch.arange(0, n_in, rate)
        idx0, idx1 = steps.long().to(torch.int32), (steps + 1).long().to(torch.int32)
        r0, r1 = F._phase_vocoder_ranges(idx0, idx1, n_in)
        assert r0.shape == r1.shape == (n_in, 2) and r0.dtype == torch.int32
        for idx, r in ((idx0, r0), (idx1, r1)):
            covered = torch.zeros(idx.numel(), dtype=torch.int32)
            for i in range(n_in):
                lo, hi = int(r[i, 0]), int(r[i, 1])
                assert 0 <= lo <= hi <= idx.numel()
                assert bool((idx[lo:hi] == i).all())
                covered[lo:hi] += 1
            inside = idx < n_in                                   # steps whose frame is one of the two appended zero frames have no gradient
            assert bool((covered[inside] == 1).all()) and bool((covered[~inside] == 0).all())


def test_forgetful_selection_finds_the_median():
    """csrc/hpss.cu finds the median of 2 m + 1 values by forgetful selection (keep m + 2 candidates, drop their minimum and
    maximum, take in the next value) on a window padded to the template size with equally many -inf and +inf.  The same
    procedure in Python against numpy's median: random windows, ties, every odd size up to 63 in each template bucket."""
    import numpy as np

    def forgetful(v):
        n = len(v)
        m = n // 2 + 2
        a = list(v[:m])
        nxt = m
        for s in range(m, 2, -1):
            for i in range(s // 2):
                if a[i] > a[s - 1 - i]:
                    a[i], a[s - 1 - i] = a[s - 1 - i], a[i]
            for i in range(1, (s + 1) // 2):
                if a[0] > a[i]:
                    a[0], a[i] = a[i], a[0]
            for i in range(s // 2, s - 1):
                if a[i] > a[s - 1]:
                    a[i], a[s - 1] = a[s - 1], a[i]
            if nxt < n:
                a[0] = v[nxt]
                nxt += 1
        assert nxt == n
        return a[1]

    rng = np.random.default_rng(5)
    for k in list(range(3, 64, 2)):
        kmax = 7 if k <= 7 else (15 if k <= 15 else (31 if k <= 31 else 63))
        fill = (kmax - k) // 2
        for trial in range(6):
            w = rng.integers(0, 9, k).astype(np.float32) if trial % 2 else rng.standard_normal(k).astype(np.float32)
            padded = [-np.inf] * fill + list(w) + [np.inf] * fill
            assert forgetful(padded) == np.median(w), (k, trial)


def test_folded_dft_and_twiddle_rotation_model():
    """stft_dft_kernel (csrc/stft.cu) evaluates X[k] = x[0] + (-1)^k x[N/2] + sum_{n=1}^{H} (x[n] + x[N-n]) cos - i (x[n] - x[N-n]) sin,
    H = (N - 1) / 2, with the twiddle exp(-2 pi i n k / N) re-read from a table every 16 steps (index n k mod N kept exactly)
    and rotated by exp(-2 pi i k / N) in between.  The same procedure in float32 numpy against numpy's rfft, even and odd sizes."""
    import numpy as np
    rng = np.random.default_rng(7)
    for n_fft in (400, 441, 6, 7, 1200):
        x = rng.standard_normal(n_fft).astype(np.float32)
        tab = np.exp(-2j * np.pi * np.arange(n_fft) / n_fft)
        tab = (tab.real.astype(np.float32), tab.imag.astype(np.float32))
        h = (n_fft - 1) // 2
        e = np.array([x[n] + x[n_fft - n] for n in range(1, h + 1)], dtype=np.float32)
        o = np.array([x[n] - x[n_fft - n] for n in range(1, h + 1)], dtype=np.float32)
        got = np.zeros(n_fft // 2 + 1, dtype=np.complex64)
        for k in range(n_fft // 2 + 1):
            re = np.float32(x[0] + ((-x[n_fft // 2] if k & 1 else x[n_fft // 2]) if n_fft % 2 == 0 else 0.0))
            im = np.float32(0)
            rot = (tab[0][k], tab[1][k])
            idx, step = k % n_fft, (16 * k) % n_fft
            for n0 in range(1, h + 1, 16):
                w = (tab[0][idx], tab[1][idx])
                for n in range(n0, min(n0 + 16, h + 1)):
                    re = np.float32(re + e[n - 1] * w[0])
                    im = np.float32(im + o[n - 1] * w[1])
                    w = (np.float32(w[0] * rot[0] - w[1] * rot[1]), np.float32(w[0] * rot[1] + w[1] * rot[0]))
                idx = (idx + step) % n_fft
            got[k] = re + 1j * (0.0 if (k == 0 or 2 * k == n_fft) else im)
        want = np.fft.rfft(x.astype(np.float64))
        assert np.abs(got - want).max() < 2e-5 * np.abs(want).max(), n_fft
