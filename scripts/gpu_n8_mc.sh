#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29528 bench.py --gpus 8 --steps 100 --warmup 3 --cpu-seconds 0.5 > gpurun_out/bench_cfg2_n8.json 2> gpurun_out/bench_cfg2_n8.err
tail -3 gpurun_out/bench_cfg2_n8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_cfg2_n8.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','with_allgather','with_peer_gather','with_multicast_gather','e2e','cfg4'):
    print(k, d.get(k))
PY
