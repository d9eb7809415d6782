set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m "gpu and not slow" -x -k "host_pipeline or cfg1" 2>&1 | tail -4
python scripts/gpu_e2e_slices.py
