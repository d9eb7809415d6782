set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m "gpu and not slow" 2>&1 | tail -4
python bench.py --steps 500 --warmup 5 --cpu-seconds 0.5 > gpurun_out/bench_cfg2.json; python -c "
import json; d=json.load(open('gpurun_out/bench_cfg2.json')); print(d['value'], d['ms_per_step'], d['kernel_ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'])"
ncu --set full --clock-control none --import-source on -k regex:stft2048 -s 4 -c 1 -o gpurun_out/prof_stft -f python bench.py --steps 2 --warmup 3 --cpu-seconds 0.1 > gpurun_out/ncu_stft.log 2>&1
