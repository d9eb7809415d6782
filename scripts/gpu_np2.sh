#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_backward.py -q -x -k "non_power or refuse or window or spectrogram or stft_backward" 2>&1 | tail -8
