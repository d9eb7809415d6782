set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_peers.py -q -x 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --workload cfg4 --steps 30 --warmup 3 --cpu-seconds 1 > gpurun_out/bench_cfg4_n2.json 2> gpurun_out/bench_cfg4_n2.err; tail -5 gpurun_out/bench_cfg4_n2.err; python - <<'PY'
import json
try:
    d = [json.loads(l) for l in open("gpurun_out/bench_cfg4_n2.json") if l.startswith("{")][-1]
    print({k: d[k] for k in ("value", "ms_per_step", "n_gpus", "scaling", "with_allgather", "with_peer_gather", "e2e") if k in d})
    print(d["config"])
except Exception as e:
    print("no bench line:", e)
PY
