#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_peers.py -q -x -rs -k "single_rank" 2>&1 | tail -15
