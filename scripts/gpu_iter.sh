# parity tests + bench + launch list (warm caches) + full captures; run under gpurun
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m "gpu and not slow" -x 2>&1 | tail -15 > gpurun_out/t_all.log; cat gpurun_out/t_all.log
python bench.py --steps 500 --warmup 5 --cpu-seconds 3 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; cat gpurun_out/bench_cfg2.json | cut -c1-2500; tail -5 gpurun_out/bench_cfg2.err
ncu --cache-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -s 6 -c 12 --csv --log-file gpurun_out/launches_warm.csv python bench.py --steps 3 --warmup 3 --cpu-seconds 0.1 > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stft2048 -s 4 -c 1 -o gpurun_out/prof_stft -f python bench.py --steps 2 --warmup 3 --cpu-seconds 0.1 > gpurun_out/ncu_stft.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:melbank -s 4 -c 1 -o gpurun_out/prof_melbank -f python bench.py --steps 2 --warmup 3 --cpu-seconds 0.1 > gpurun_out/ncu_melbank.log 2>&1
ls gpurun_out
