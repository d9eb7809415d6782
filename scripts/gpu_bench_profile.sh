# bench + launch list + full ncu captures of the two pipeline kernels (run under gpurun)
set -x
mkdir -p gpurun_out
python __graft_entry__.py --smoke 2>&1 | tail -3
python bench.py --steps 50 --warmup 5 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; tail -c 3000 gpurun_out/bench_cfg2.json; tail -5 gpurun_out/bench_cfg2.err
python bench.py --steps 10 --warmup 3 --workload cfg3 --cpu-seconds 5 > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; tail -c 3000 gpurun_out/bench_cfg3.json; tail -5 gpurun_out/bench_cfg3.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --cpu-seconds 0.2 > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stft2048 -s 4 -c 1 -o gpurun_out/prof_stft -f python bench.py --steps 2 --warmup 3 --cpu-seconds 0.2 > gpurun_out/ncu_stft.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:melbank -s 4 -c 1 -o gpurun_out/prof_melbank -f python bench.py --steps 2 --warmup 3 --cpu-seconds 0.2 > gpurun_out/ncu_melbank.log 2>&1
ls -la gpurun_out
