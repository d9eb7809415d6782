# two GPUs: peer-gather parity test + the N=2 bench (NCCL all-gather leg next to the kernel-side gather)
set -x
mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -8
timeout 400 python -m pytest tests/test_gpu_peers.py -q -x 2>&1 | tail -15
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 300 --warmup 5 --cpu-seconds 1 > gpurun_out/bench_cfg2_n2.json 2> gpurun_out/bench_cfg2_n2.err; tail -5 gpurun_out/bench_cfg2_n2.err; python - <<'PY'
import json
try:
    d = [json.loads(l) for l in open("gpurun_out/bench_cfg2_n2.json") if l.startswith("{")][-1]
    print({k: d[k] for k in ("value", "ms_per_step", "n_gpus", "with_allgather", "with_peer_gather") if k in d})
except Exception as e:
    print("no bench line:", e)
PY
