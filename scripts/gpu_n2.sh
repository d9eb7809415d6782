set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 500 --warmup 5 --cpu-seconds 1 > gpurun_out/bench_cfg2_n2.json 2> gpurun_out/bench_n2.err; cut -c1-700 gpurun_out/bench_cfg2_n2.json; tail -3 gpurun_out/bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference --steps 5 --warmup 1 | cut -c1-200
