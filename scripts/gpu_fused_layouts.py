"""A/B of the one-kernel mel path's output layouts and of the two-kernel path, config 2 shape (CUDA events,
one C-ABI call per step into a pre-allocated output)."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchaudio_contrib_b200 as tac

dev = torch.device("cuda")
fb = tac.MelFilterbank(num_freqs=1025, num_mels=128, sample_rate=16000).get_filterbank()
xs = [torch.randn(64, 1, 160000, device=dev) for _ in range(7)]


def timed(prep, n=500):
    out = prep.empty_output()
    for i in range(5):
        prep(xs[i % 7], out)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n):
        prep(xs[i % 7], out)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for layout in ("contiguous", "reference"):
    for db in (False, True):
        ms = timed(tac.PreparedMelspectrogram((64, 1, 160000), dev, fb, 2048, 512, to_db=db, layout=layout))
        print("fused layout=%-10s to_db=%d  %.4f ms/step  %.3e frames/s" % (layout, db, ms, 20032 / ms * 1e3))
os.environ["TAC_MELSPEC_FUSED"] = "0"
ms = timed(tac.PreparedMelspectrogram((64, 1, 160000), dev, fb, 2048, 512))
print("two-kernel path                    %.4f ms/step  %.3e frames/s" % (ms, 20032 / ms * 1e3))
