#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m "gpu and not slow" -x 2>&1 | tail -8
python __graft_entry__.py --smoke 2>&1 | tail -1
python bench.py --steps 200 --warmup 5 --cpu-seconds 2 > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err; tail -2 gpurun_out/bench_r02.err; cut -c1-400 gpurun_out/bench_r02.json
