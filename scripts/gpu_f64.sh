#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "float64 or non_power" 2>&1 | tail -15
