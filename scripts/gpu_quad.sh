#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_peers.py -q -x -k "pair or fused or mel or peer or full_size or graph" 2>&1 | tail -5
python scripts/gpu_time_variant.py
