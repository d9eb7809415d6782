set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m "gpu and not slow" -x -k "mel or fused or host_pipeline or stft" 2>&1 | tail -5
python scripts/gpu_fused_layouts.py
