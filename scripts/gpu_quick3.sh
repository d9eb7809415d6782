set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m "gpu and not slow" 2>&1 | tail -12
python scripts/gpu_time_ops.py 2>&1 | grep -E "fft|unfused"
