"""Event-timed throughput of the non-headline entry points (frames/s) on (64,1,160000)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchaudio_contrib_b200 as tac

def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

x = torch.randn(64, 1, 160000, device="cuda")
with torch.no_grad():
    for fft in (256, 512, 1024, 2048, 4096):
        hop = fft // 4
        frames = 64 * (1 + 160000 // hop)
        st = tac.STFT(fft, hop).cuda()
        sp = tac.Spectrogram(fft, hop, power=2.0).cuda()
        mel = tac.Melspectrogram(num_mels=128, sample_rate=16000, fft_length=fft, hop_length=hop).cuda()
        t1, t2, t3 = timeit(lambda: st(x)), timeit(lambda: sp(x)), timeit(lambda: mel(x))
        print("fft %4d hop %4d frames %7d | stft %.3f ms %.2e f/s | spectrogram %.3f ms %.2e f/s | mel %.3f ms %.2e f/s"
              % (fft, hop, frames, t1, frames / t1 * 1e3, t2, frames / t2 * 1e3, t3, frames / t3 * 1e3))
    st = tac.STFT(2048, 512).cuda()
    z = st(x)
    cn = tac.ComplexNorm(2.0)
    fb = tac.ApplyFilterbank(tac.MelFilterbank(1025, 128, sample_rate=16000).get_filterbank()).cuda()
    p = cn(z)
    print("complex_norm %.3f ms  apply_filterbank(public) %.3f ms  amplitude_to_db %.3f ms"
          % (timeit(lambda: cn(z)), timeit(lambda: fb(p)), timeit(lambda: tac.amplitude_to_db(p))))
    chain = torch.nn.Sequential(st, cn, fb)
    print("unfused chain (plain nn.Sequential) %.3f ms" % timeit(lambda: chain(x)))
    # round 2: sizes that are not a power of two (direct-DFT kernel), the dense-filterbank path, double tensors
    for fft, hop in ((400, 160), (1200, 300), (441, 110)):
        frames = 64 * (1 + 160000 // hop)
        sp = tac.Spectrogram(fft, hop, power=2.0).cuda()
        mel = tac.Melspectrogram(num_mels=80, sample_rate=16000, fft_length=fft, hop_length=hop).cuda()
        t2, t3 = timeit(lambda: sp(x), 5), timeit(lambda: mel(x), 5)
        print("fft %4d hop %4d frames %7d | spectrogram %.3f ms %.2e f/s | mel(80) %.3f ms %.2e f/s   (direct DFT)"
              % (fft, hop, frames, t2, frames / t2 * 1e3, t3, frames / t3 * 1e3))
    dense = torch.randn(1025, 128, device="cuda").abs()
    t = timeit(lambda: tac.functional.melspectrogram(x, dense, 2048, 512))
    print("dense 1025 x 128 filterbank, fft 2048 (K1 -> K2, tcgen05): %.3f ms %.2e f/s" % (t, 64 * 313 / t * 1e3))
    xd = x[:8].double()
    t = timeit(lambda: tac.stft(xd, 512, 128), 3)
    print("float64 stft 512/128 on (8,1,160000): %.3f ms" % t)
    # CUDA-graph replay of the small, launch-bound case (BASELINE config 1 shape through the mel chain)
    fb = tac.MelFilterbank(num_freqs=257, num_mels=64, sample_rate=16000).get_filterbank()
    xs = torch.randn(1, 1, 16000, device="cuda")
    prep = tac.PreparedMelspectrogram(xs.shape, "cuda", fb, 512, 128, to_db=True)
    o = prep.empty_output()
    prep(xs, o); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20):
            prep(xs, o)
    te = timeit(lambda: [prep(xs, o) for _ in range(20)], 20) / 20
    tg = timeit(g.replay, 20) / 20
    print("(1,1,16000) fft 512 mel+dB: eager call %.2f us, inside a replayed CUDA graph %.2f us per launch" % (te * 1e3, tg * 1e3))
