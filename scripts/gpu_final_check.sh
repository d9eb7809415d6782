# last check of the round: every GPU test but the 2^32 sweep, smoke, one short default bench line
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m "gpu and not slow" 2>&1 | tail -6
python __graft_entry__.py --smoke 2>&1 | tail -1
python bench.py --steps 200 --warmup 5 --cpu-seconds 2 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -2 gpurun_out/bench_final.err; cut -c1-260 gpurun_out/bench_final.json
