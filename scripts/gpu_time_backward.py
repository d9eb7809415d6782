"""Forward + backward timing of the mel chain at BASELINE config 2 ((64,1,160000), fft 2048 / hop 512, 128 mels)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchaudio_contrib_b200 as tac

dev = torch.device("cuda")
model = tac.Melspectrogram(num_mels=128, sample_rate=16000, fft_length=2048, hop_length=512).to(dev)
model_db = tac.Sequential(*model, tac.AmplitudeToDb()).to(dev)
xs = [torch.randn(64, 1, 160000, device=dev, requires_grad=True) for _ in range(4)]


def timed(fn, n=20):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def fwd_bwd(m):
    def run(i):
        x = xs[i % 4]
        x.grad = None
        y = m(x)
        y.backward(gy[0])
    return run


with torch.no_grad():
    y0 = model(xs[0])
gy = [torch.randn_like(y0)]
lib = tac._cabi.lib()
print("forward only (no grad)            %.3f ms" % timed(lambda i: model(xs[i % 4].detach())))
print("forward + backward, mel           %.3f ms" % timed(fwd_bwd(model)))
print("forward + backward, mel + dB      %.3f ms" % timed(fwd_bwd(model_db)))
spec = tac.Spectrogram(fft_length=2048, hop_length=512, power=2.0).to(dev)
with torch.no_grad():
    s0 = spec(xs[0])
gy.append(torch.randn_like(s0))
def spec_run(i):
    x = xs[i % 4]
    x.grad = None
    spec(x).backward(gy[1])
print("forward + backward, Spectrogram   %.3f ms" % timed(spec_run))
import ctypes
for name, m in (("mel", model), ("mel+dB", model_db)):
    lib.tac_profile_enable(1)
    fwd_bwd(m)(0)
    torch.cuda.synchronize()
    ms = (ctypes.c_double * 4)()
    n = (ctypes.c_int64 * 4)()
    lib.tac_profile_read(ms, n)
    lib.tac_profile_enable(0)
    print("one fwd+bwd (%s) by kernel kind: stft-family %.3f ms (%d launches), filterbank %.3f ms (%d), pointwise %.3f ms (%d)"
          % (name, ms[0], n[0], ms[1], n[1], ms[3], n[3]))
print("upstream gradient strides (mel):", gy[0].stride())
