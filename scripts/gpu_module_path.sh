#!/bin/bash
timeout 1500 python -m pytest tests -q -x -m "gpu and not slow" 2>&1 | tail -4
python bench.py --steps 500 --warmup 5 --cpu-seconds 0.5 --skip-extras 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('call', d['ms_per_step'], 'module', d['module_ms_per_step'], d['module_matches_call'], 'traffic', d['roofline']['traffic'])"
