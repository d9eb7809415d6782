#!/bin/bash
timeout 600 python -m pytest tests/test_backward.py -q -x -k "phase_vocoder" 2>&1 | tail -15
