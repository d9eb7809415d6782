"""A/B of the one-kernel mel variants (0 pair, 1 one frame per warp, 2 pair + tcgen05 pass): agreement on assorted shapes,
then event timing at configs 2 and 3."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchaudio_contrib_b200 as tac
from torchaudio_contrib_b200 import _cabi

lib = _cabi.lib()


def run(variant, mod, x):
    lib.tac_mel_kernel_variant(variant)
    with torch.no_grad():
        y = mod(x).contiguous()
    torch.cuda.synchronize()
    return y


def timeit(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


torch.manual_seed(0)
ok = True
cases = [((3, 1, 16000), 16000, "reflect", False), ((2, 2, 48001), 48000, "reflect", True), ((5, 1, 4096), 16000, "constant", False),
         ((1, 1, 2049), 22050, "replicate", True), ((7, 1, 160000), 16000, "reflect", False), ((4, 1, 33333), 16000, "circular", False),
         ((1, 3, 2048 * 3 + 17), 8000, "reflect", True)]
for shape, sr, pad_mode, db in cases:
    x = torch.randn(*shape, device="cuda")
    mel = tac.Melspectrogram(num_mels=128, sample_rate=sr, fft_length=2048, hop_length=512, pad_mode=pad_mode).cuda()
    mod = tac.Sequential(*mel, tac.AmplitudeToDb()).cuda() if db else mel
    y0, y1, y2 = run(1, mod, x), run(0, mod, x), run(2, mod, x)
    for label, ya in (("pair", y1), ("pair+tc", y2)):
        err = (y0 - ya).abs().max().item() if db else ((y0 - ya).abs() / y0.abs().clamp_min(1e-30)).max().item()
        same = err < (1e-4 if db else 2e-5)
        ok &= same
        print("shape %-18s sr %5d pad %-9s db %d: %-7s ~ single: %s (max %s diff %.3g, out %s)"
              % (shape, sr, pad_mode, db, label, same, "dB" if db else "rel", err, tuple(y0.shape)))
# unaligned view (bulk copy impossible -> gather path)
xb = torch.randn(3, 1, 20001, device="cuda")[:, :, 1:]
mel = tac.Melspectrogram(num_mels=128, sample_rate=16000, fft_length=2048, hop_length=512).cuda()
y0, y1, y2 = run(1, mel, xb), run(0, mel, xb), run(2, mel, xb)
for ya in (y1, y2):
    err = ((y0 - ya).abs() / y0.abs().clamp_min(1e-30)).max().item()
    print("unaligned view: max rel diff %.3g" % err); ok &= err < 2e-5
print("ALL CLOSE" if ok else "MISMATCH")

for name, shape, sr, db in (("cfg2", (64, 1, 160000), 16000, False), ("cfg3", (256, 2, 480000), 48000, True)):
    x = torch.randn(*shape, device="cuda")
    mel = tac.Melspectrogram(num_mels=128, sample_rate=sr, fft_length=2048, hop_length=512).cuda()
    mod = tac.Sequential(*mel, tac.AmplitudeToDb()).cuda() if db else mel
    frames = shape[0] * shape[1] * (1 + shape[2] // 512)
    with torch.no_grad():
        for variant, label in ((1, "single "), (0, "pair   "), (2, "pair+tc")):
            lib.tac_mel_kernel_variant(variant)
            t = timeit(lambda: mod(x), 50 if name == "cfg2" else 10)
            print("%s %s: %.4f ms per step, %.3e frames/s" % (name, label, t, frames / t * 1e3))
lib.tac_mel_kernel_variant(0)
