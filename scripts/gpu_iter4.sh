set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m "gpu and not slow" -x 2>&1 | tail -4
python scripts/gpu_fused_layouts.py
