"""e2e (host buffers) timing of the cfg2 step against the host pipeline's slice size (TAC_HOST_SLICE_MB)."""
import os, sys, time, subprocess
if len(sys.argv) > 1:
    if sys.argv[1] != "default":
        os.environ["TAC_HOST_SLICE_MB"] = sys.argv[1]
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch
    import torchaudio_contrib_b200 as tac
    fb = tac.MelFilterbank(num_freqs=1025, num_mels=128, sample_rate=16000).get_filterbank()
    hp = tac.HostPipeline(2048, 512, power=2.0, filterbank=fb, to_db=False, device="cuda")
    xs = [torch.randn(64, 1, 160000).pin_memory() for _ in range(2)]
    out = torch.empty(64, 1, 128, 313).pin_memory()
    for i in range(5):
        hp(xs[i % 2], out=out)
    torch.cuda.synchronize()
    best = []
    for rep in range(3):
        t0 = time.perf_counter()
        for i in range(40):
            hp(xs[i % 2], out=out)
        torch.cuda.synchronize()
        best.append((time.perf_counter() - t0) / 40 * 1e3)
    # plain copies for scale: the whole input in one cudaMemcpyAsync, and in 8 concurrent pieces
    d = torch.empty(64, 1, 160000, device="cuda")
    def one():
        d.copy_(xs[0], non_blocking=True)
    streams = [torch.cuda.Stream() for _ in range(8)]
    def eight():
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                d[8 * i:8 * i + 8].copy_(xs[0][8 * i:8 * i + 8], non_blocking=True)
    res = []
    for fn in (one, eight):
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(20):
            fn()
        torch.cuda.synchronize()
        res.append(40.96e6 * 20 / (time.perf_counter() - t0) / 1e9)
    print("slice %7s MB  e2e %.3f ms/step   (plain H2D of the input: 1 copy %.1f GB/s, 8 concurrent %.1f GB/s)"
          % (sys.argv[1], min(best), res[0], res[1]))
else:
    for v in ("default", "1", "2", "3", "5", "8", "14"):
        subprocess.run([sys.executable, __file__, v])
