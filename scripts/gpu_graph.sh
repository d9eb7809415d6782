#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "cuda_graph" 2>&1 | tail -8
python scripts/gpu_time_ops.py 2>&1 | tail -3
