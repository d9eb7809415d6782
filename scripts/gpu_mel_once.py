"""A handful of one-kernel mel calls at BASELINE config 2 (or `cfg3`): the short command ncu wraps."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchaudio_contrib_b200 as tac

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 6
shape, sr, db = ((64, 1, 160000), 16000, False) if cfg == "cfg2" else ((256, 2, 480000), 48000, True)
mods = list(tac.Melspectrogram(num_mels=128, sample_rate=sr, fft_length=2048, hop_length=512))
if db:
    mods.append(tac.AmplitudeToDb())
m = tac.Sequential(*mods).cuda()
xs = [torch.randn(*shape, device="cuda") for _ in range(2)]
with torch.no_grad():
    for i in range(n):
        y = m(xs[i % 2])
torch.cuda.synchronize()
print(cfg, tuple(y.shape), float(y.mean()))
