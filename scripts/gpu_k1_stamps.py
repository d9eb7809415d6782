"""Wall-clock stamps of every warp of the one-kernel mel path (timing build: scripts/build_trace_lib.sh)."""
import os, sys, ctypes, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["TAC_K1_TRACE"] = "1"
os.environ["TAC_B200_LIB"] = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                           "torchaudio_contrib_b200", "lib", "trace", "libtac_b200_trace.so")
import torchaudio_contrib_b200 as tac
m = tac.Melspectrogram(num_mels=128, sample_rate=16000, fft_length=2048, hop_length=512).cuda()
xs = [torch.randn(64, 1, 160000, device="cuda") for _ in range(7)]
for i in range(6):
    y = m(xs[i])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
y = m(xs[6])
e1.record()
torch.cuda.synchronize()
print("event-timed launch: %.1f us" % (e0.elapsed_time(e1) * 1e3))
tac._cabi.lib().tac_debug_dump_k1_trace.restype = ctypes.c_int
tac._cabi.lib().tac_debug_dump_k1_trace()
