# eight GPUs: N=8 bench (cfg2 per rank) with both gather legs
set -x
mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 100 --warmup 3 --cpu-seconds 0.5 > gpurun_out/bench_cfg2_n8.json 2> gpurun_out/bench_cfg2_n8.err; tail -3 gpurun_out/bench_cfg2_n8.err; python - <<'PY'
import json
try:
    d = [json.loads(l) for l in open("gpurun_out/bench_cfg2_n8.json") if l.startswith("{")][-1]
    print({k: d[k] for k in ("value", "ms_per_step", "n_gpus", "with_allgather", "with_peer_gather", "e2e") if k in d})
except Exception as e:
    print("no bench line:", e)
PY
