# four GPUs: N=4 bench with both gather legs
set -x
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 200 --warmup 5 --cpu-seconds 1 > gpurun_out/bench_cfg2_n4.json 2> gpurun_out/bench_cfg2_n4.err; tail -5 gpurun_out/bench_cfg2_n4.err; python - <<'PY'
import json
try:
    d = [json.loads(l) for l in open("gpurun_out/bench_cfg2_n4.json") if l.startswith("{")][-1]
    print({k: d[k] for k in ("value", "ms_per_step", "n_gpus", "with_allgather", "with_peer_gather") if k in d})
except Exception as e:
    print("no bench line:", e)
PY
