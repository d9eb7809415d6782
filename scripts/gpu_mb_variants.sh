set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m "gpu and not slow" 2>&1 | tail -5
for v in 0 1 2 3 4 7; do
  echo "variant $v"; TAC_MB_DEBUG=$v python bench.py --steps 200 --warmup 5 --cpu-seconds 0.1 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['kernel_ms_per_step'])"
done
