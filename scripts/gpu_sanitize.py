"""Small invocations of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchaudio_contrib_b200 as tac

torch.manual_seed(0)
dev = "cuda"
x = torch.randn(3, 2, 9000, device=dev)
mel = tac.Melspectrogram(num_mels=128, sample_rate=16000, fft_length=2048, hop_length=512).to(dev)
db = tac.Sequential(*mel, tac.AmplitudeToDb()).to(dev)
y = db(x)                                                                         # one-kernel path, edge frames included
y2 = tac.Melspectrogram(num_mels=64, sample_rate=16000, fft_length=1024, hop_length=256).to(dev)(x)   # warp kernel + tcgen05 filterbank
y3 = tac.Spectrogram(fft_length=512, hop_length=128).to(dev)(x)
y4 = tac.stft(x, 2048, 512)
y5 = tac.stft(x, 4096, 1024)                                                      # generic kernel
odd = torch.randn(2, 1, 7001, device=dev)
y6 = mel(odd)                                                                     # unaligned rows: gather path
codes = tac.mu_law_encoding(torch.rand(100000, device=dev) * 2 - 1)
back = tac.mu_law_decoding(codes)
z = tac.stft(x, 512, 128)
pv = tac.phase_vocoder(z, 1.3, torch.linspace(0, 3.14159265 * 128, 257, device=dev)[..., None])
xg = x.clone().requires_grad_(True)
db(xg).sum().backward()                                                           # adjoint kernels
hp = tac.HostPipeline(2048, 512, power=2.0, filterbank=mel[2].filterbank, to_db=False, device=dev)
out = hp(torch.randn(5, 1, 20000).pin_memory())
torch.cuda.synchronize()
print("sanitize run ok", float(y.sum()), float(xg.grad.abs().sum()), float(out.sum()))
