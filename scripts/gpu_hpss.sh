#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "hpss" 2>&1 | tail -12
python - <<'PY'
import torch, sys
sys.path.insert(0,'.')
from torchaudio_contrib_b200.beta_hpss import hpss
x = torch.rand(64, 1, 1025, 313, device="cuda")
for k in (31, 17):
    for _ in range(2): hpss(x, k)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): hpss(x, k)
    b.record(); torch.cuda.synchronize()
    print("hpss kernel_size %d on (64,1,1025,313): %.3f ms" % (k, a.elapsed_time(b) / 5))
PY
