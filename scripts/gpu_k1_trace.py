import os, sys, ctypes, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["TAC_K1_TRACE"] = "1"
import torchaudio_contrib_b200 as tac
m = tac.Melspectrogram(num_mels=128, sample_rate=16000, fft_length=2048, hop_length=512).cuda()
x = torch.randn(64, 1, 160000, device="cuda")
y = m(x)
torch.cuda.synchronize()
lib = ctypes.CDLL(tac._cabi.LIB_PATH)
lib.tac_debug_dump_k1_trace()
