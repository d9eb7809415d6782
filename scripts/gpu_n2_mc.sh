#!/bin/bash
# two GPUs: peer + multicast gather tests, then the N = 2 bench line (all-gather, unicast peer stores, multicast stores)
mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -8
timeout 400 python -m pytest tests/test_gpu_peers.py -q -x -rs 2>&1 | tail -12
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 100 --warmup 3 --cpu-seconds 0.5 > gpurun_out/bench_cfg2_n2.json 2> gpurun_out/bench_cfg2_n2.err
tail -3 gpurun_out/bench_cfg2_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_cfg2_n2.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','with_allgather','with_peer_gather','with_multicast_gather'):
    print(k, d.get(k))
PY
