set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m "gpu and not slow" 2>&1 | tail -5
python scripts/gpu_mb_trace.py 2>&1 | tail -5 | cut -c1-900
python bench.py --steps 500 --warmup 5 --cpu-seconds 0.5 > gpurun_out/bench_cfg2.json; python -c "
import json; d=json.load(open('gpurun_out/bench_cfg2.json')); print(d['value'], d['ms_per_step'], d['kernel_ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'])"
