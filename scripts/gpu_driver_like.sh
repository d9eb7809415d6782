#!/bin/bash
# what the driver runs at round end: the whole -m gpu suite (incl. the 2^32 mu-law sweep), smoke, the default bench line
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4 ) 2>&1 | tail -8
python __graft_entry__.py --smoke 2>&1 | tail -1
( time python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2>&1 | tail -3
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value','ms_per_step','steps','gpu_launches')}); print(d['roofline']); print(d['e2e']['ms_per_step'], d['e2e']['value']); print(d['clocks'])
PY
