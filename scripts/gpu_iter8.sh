set -x
timeout 600 python -m pytest tests/test_backward.py -q -m gpu -x 2>&1 | tail -5
python scripts/gpu_time_backward.py
