# all GPU parity tests (not the 2^32 sweep) + layout A/B + bench
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m "gpu and not slow" 2>&1 | tail -25 > gpurun_out/t_all.log; cat gpurun_out/t_all.log
python scripts/gpu_fused_layouts.py > gpurun_out/fused_layouts.txt 2>&1; cat gpurun_out/fused_layouts.txt
python bench.py --steps 500 --warmup 5 --cpu-seconds 2 > gpurun_out/bench_fused.json 2> gpurun_out/bench_fused.err; cut -c1-600 gpurun_out/bench_fused.json; tail -3 gpurun_out/bench_fused.err
