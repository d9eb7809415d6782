#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "non_power or float64 or stft" 2>&1 | tail -6
python - <<'PY'
import torch, sys
sys.path.insert(0,'.')
import torchaudio_contrib_b200 as tac
x = torch.randn(64, 1, 160000, device="cuda")
def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
with torch.no_grad():
    for fft, hop in ((400, 160), (1200, 300), (441, 110), (3000, 750)):
        frames = 64 * (1 + 160000 // hop)
        sp = tac.Spectrogram(fft, hop, power=2.0).cuda()
        t = timeit(lambda: sp(x))
        print("fft %4d hop %4d frames %7d | spectrogram %.3f ms %.2e f/s" % (fft, hop, frames, t, frames / t * 1e3))
PY
