#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_peers.py -q -x -k "pair or fused or mel or peer or full_size" 2>&1 | tail -5
