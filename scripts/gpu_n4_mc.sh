#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 4 --steps 100 --warmup 3 --cpu-seconds 0.5 > gpurun_out/bench_cfg2_n4.json 2> gpurun_out/bench_cfg2_n4.err
tail -3 gpurun_out/bench_cfg2_n4.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_cfg2_n4.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','with_allgather','with_peer_gather','with_multicast_gather','e2e','cfg4','cfg3'):
    print(k, json.dumps(d.get(k))[:400])
PY
