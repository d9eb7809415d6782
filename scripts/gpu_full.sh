# full evidence run: all GPU tests (incl. the 2^32 mu-law sweep), benches, launch list, full captures
set -x
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep gpurun_out/launches*.csv
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -6 > gpurun_out/t_gpu_all.log; cat gpurun_out/t_gpu_all.log
python bench.py --steps 1000 --warmup 5 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; cut -c1-1200 gpurun_out/bench_cfg2.json
python scripts/gpu_time_ops.py > gpurun_out/time_ops.txt 2>&1; cat gpurun_out/time_ops.txt
python bench.py --steps 20 --warmup 3 --workload cfg3 --cpu-seconds 4 > gpurun_out/bench_cfg3.json 2>/dev/null; cut -c1-400 gpurun_out/bench_cfg3.json
python bench.py --workload mulaw --steps 20 > gpurun_out/bench_mulaw.json 2>/dev/null; cut -c1-1500 gpurun_out/bench_mulaw.json
python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/bench_reference.json 2>/dev/null; cut -c1-300 gpurun_out/bench_reference.json
ncu --cache-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -s 6 -c 16 --csv --log-file gpurun_out/launches_warm.csv python bench.py --steps 4 --warmup 3 --cpu-seconds 0.1 > gpurun_out/ncu_launches.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 6 -c 16 --csv --log-file gpurun_out/launches.csv python bench.py --steps 4 --warmup 3 --cpu-seconds 0.1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:stft2048 -s 4 -c 1 -o gpurun_out/prof_stft -f python bench.py --steps 2 --warmup 3 --cpu-seconds 0.1 > gpurun_out/ncu_stft.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:melbank -s 4 -c 1 -o gpurun_out/prof_melbank -f python bench.py --steps 2 --warmup 3 --cpu-seconds 0.1 > gpurun_out/ncu_melbank.log 2>&1
ncu --set full --clock-control none -k regex:mulaw_encode -s 1 -c 1 -o gpurun_out/prof_mulaw -f python bench.py --workload mulaw --steps 2 --warmup 3 > gpurun_out/ncu_mulaw.log 2>&1
ls gpurun_out
