# quick loop for the one-kernel mel path: parity, timing, per-warp stamps (timing build)
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m "gpu and not slow" -x -k "mel or fused or host_pipeline or stft or spectrogram" 2>&1 | tail -4
python scripts/gpu_fused_layouts.py
python scripts/gpu_k1_stamps.py 2>&1 | grep -v "^k1 trace\|^first data"
