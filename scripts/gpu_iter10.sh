set -x
timeout 900 python -m pytest tests/test_backward.py -q -m gpu -x 2>&1 | tail -8
python scripts/gpu_time_backward.py
