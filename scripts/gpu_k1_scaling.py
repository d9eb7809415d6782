"""K1 / K2 launch time versus number of frames (whole waves of 2368 frames)."""
import os, sys, ctypes, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["TAC_MELSPEC_SLICE_ROWS"] = "100000000"
import torchaudio_contrib_b200 as tac
lib = tac._cabi.lib()
m = tac.Melspectrogram(num_mels=128, sample_rate=16000, fft_length=2048, hop_length=512).cuda()
T = int(os.environ.get("T", 160000))
FPS = 1 + T // 512
for n_seq in (8, 16, 32, 64, 128, 256):
    xs = [torch.randn(n_seq, 1, T, device="cuda") for _ in range(max(2, 400 // n_seq))]
    with torch.no_grad():
        for i in range(4): m(xs[i % len(xs)])
        torch.cuda.synchronize()
        lib.tac_profile_enable(1)
        reps = 20
        for i in range(reps): m(xs[i % len(xs)])
        torch.cuda.synchronize()
        ms = (ctypes.c_double * 4)(); cnt = (ctypes.c_int64 * 4)()
        lib.tac_profile_read(ms, cnt); lib.tac_profile_enable(0)
    frames = n_seq * FPS
    print("frames %7d (%2d waves): K1 %.1f us (%.2e f/s)  K2 %.1f us" % (frames, frames // 2368, 1e3 * ms[0] / cnt[0], frames / (ms[0] / cnt[0]) * 1e3, 1e3 * ms[1] / cnt[1]))
