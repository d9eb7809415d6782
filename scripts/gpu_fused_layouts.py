"""A/B of the one-kernel mel path's output layouts and of the two-kernel path, config 2 shape (CUDA events)."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchaudio_contrib_b200 as tac
from torchaudio_contrib_b200 import functional as F

dev = torch.device("cuda")
fb = tac.MelFilterbank(num_freqs=1025, num_mels=128, sample_rate=16000).get_filterbank().to(dev)
xs = [torch.randn(64, 1, 160000, device=dev) for _ in range(7)]
cache = {}


def timed(fn, n=300):
    for i in range(5):
        fn(xs[i % 7])
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n):
        fn(xs[i % 7])
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for layout in ("contiguous", "reference"):
    for db in (False, True):
        ms = timed(lambda x: F.melspectrogram(x, fb, 2048, 512, layout=layout, to_db=db, _cache=cache))
        print("fused layout=%-10s to_db=%d  %.4f ms/step  %.3e frames/s" % (layout, db, ms, 20032 / ms * 1e3))
os.environ["TAC_MELSPEC_FUSED"] = "0"
ms = timed(lambda x: F.melspectrogram(x, fb, 2048, 512, _cache=cache))
print("two-kernel path                    %.4f ms/step  %.3e frames/s" % (ms, 20032 / ms * 1e3))
