#!/bin/bash
timeout 600 python -m pytest tests/test_backward.py -q -x -k "window or constants" 2>&1 | tail -25
