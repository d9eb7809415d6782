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
# round 2: tensor-core variant, dense filterbank (tcgen05 kernel), 4096 / non-power-of-two sizes, double path, new adjoints,
# single-rank peer and multicast gathers (allocation, PEERS stores, flag barriers)
lib = tac._cabi.lib()
lib.tac_mel_kernel_variant(2)
y7 = db(x)
lib.tac_mel_kernel_variant(1)
y8 = db(x)
lib.tac_mel_kernel_variant(0)
dense = tac.functional.melspectrogram(x, torch.randn(1025, 40, device=dev), 2048, 512)       # K1 -> K2 (tcgen05 filterbank)
y9 = tac.Melspectrogram(num_mels=64, sample_rate=16000, fft_length=4096, hop_length=1024).to(dev)(x)
y10 = tac.Melspectrogram(num_mels=40, sample_rate=16000, fft_length=400, hop_length=160).to(dev)(x)   # direct-DFT kernel -> K2
y11 = tac.stft(x.double(), 400, 160)
y12 = tac.functional.melspectrogram(x.double(), tac.MelFilterbank(num_freqs=257, num_mels=40, sample_rate=16000).get_filterbank().double().to(dev), 512, 128, to_db=True)
zg = z.clone().requires_grad_(True)
tac.phase_vocoder(zg, 0.8, torch.linspace(0, 3.14159265 * 128, 257, device=dev)[..., None]).sum().backward()
wg = torch.hann_window(512, device=dev).requires_grad_(True)
tac.stft(x, 512, 128, window=wg).sum().backward()
xs = x[:1, :1, :4000].clone().requires_grad_(True)
tac.spectrogram(xs, 400, 160, power=1.0).sum().backward()                           # direct-DFT kernel and its adjoint
from torchaudio_contrib_b200.beta_hpss import hpss
hp_out = hpss(tac.Spectrogram(512, 128).to(dev)(x), 31)                              # median filters
import torch.distributed as dist
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29577")
dist.init_process_group("gloo", rank=0, world_size=1)
from torchaudio_contrib_b200.distributed import PeerGatheredOutput
prep = tac.PreparedMelspectrogram(x.shape, torch.device(dev, 0), mel[2].filterbank, 2048, 512, to_db=True)
buf = PeerGatheredOutput(prep.out_shape, torch.device(dev, 0))
prep.gather_into(x, buf); buf.wait(); buf.close()
# (the multicast gather needs two devices: tests/test_gpu_peers.py::test_multicast_gather_two_gpus)
dist.destroy_process_group()
hp = tac.HostPipeline(2048, 512, power=2.0, filterbank=mel[2].filterbank, to_db=False, device=dev)
out = hp(torch.randn(5, 1, 20000).pin_memory())
torch.cuda.synchronize()
print("sanitize run ok", float(y.sum()), float(xg.grad.abs().sum()), float(out.sum()))
