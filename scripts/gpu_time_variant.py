"""Event-timed one-kernel mel step at configs 2 and 3 for whatever library TAC_B200_LIB names (scripts/build_variant.sh)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchaudio_contrib_b200 as tac

def timeit(fn, n):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

tag = os.path.basename(os.environ.get("TAC_B200_LIB", "default"))
out = []
for name, shape, sr, db, n in (("cfg2", (64, 1, 160000), 16000, False, 100), ("cfg3", (256, 2, 480000), 48000, True, 10)):
    fb = tac.MelFilterbank(num_freqs=1025, num_mels=128, sample_rate=sr).get_filterbank()
    prep = tac.PreparedMelspectrogram(shape, "cuda", fb, 2048, 512, power=2.0, to_db=db)
    xs = [torch.randn(*shape, device="cuda") for _ in range(4 if name == "cfg2" else 2)]
    o = prep.empty_output()
    i = [0]
    def step():
        prep(xs[i[0] % len(xs)], o); i[0] += 1
    t = timeit(step, n)
    frames = shape[0] * shape[1] * (1 + shape[2] // 512)
    out.append("%s %.4f ms %.3e f/s" % (name, t, frames / t * 1e3))
print("%-28s %s" % (tag, " | ".join(out)))
