set -x
mkdir -p gpurun_out
python bench.py --steps 300 --warmup 5 --cpu-seconds 2 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -3 gpurun_out/bench_quick.err
python - <<'PY'
import json
d = [json.loads(l) for l in open("gpurun_out/bench_quick.json") if l.startswith("{")][-1]
print({k: d[k] for k in ("value", "ms_per_step", "e2e", "reference_on_gpu", "cpu_baseline", "roofline")})
PY
python scripts/gpu_time_backward.py
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python scripts/gpu_sanitize.py > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/sanitize_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python scripts/gpu_sanitize.py > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/sanitize_racecheck.log
