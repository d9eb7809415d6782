# evidence for the one-kernel mel path: parity tests, benches, launch lists, one full capture
set -x
mkdir -p gpurun_out; rm -f gpurun_out/launches*.csv
timeout 1200 python -m pytest tests -q -m "gpu and not slow" 2>&1 | tail -25 > gpurun_out/t_all.log; cat gpurun_out/t_all.log
python bench.py --steps 1000 --warmup 5 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; cut -c1-700 gpurun_out/bench_cfg2.json; tail -3 gpurun_out/bench_cfg2.err
python bench.py --steps 20 --warmup 3 --workload cfg3 --cpu-seconds 4 > gpurun_out/bench_cfg3.json 2>/dev/null; cut -c1-400 gpurun_out/bench_cfg3.json
python scripts/gpu_fused_layouts.py > gpurun_out/fused_layouts.txt 2>&1; cat gpurun_out/fused_layouts.txt
ncu --cache-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -s 6 -c 16 --csv --log-file gpurun_out/launches_warm.csv python bench.py --steps 4 --warmup 3 --cpu-seconds 0.1 > gpurun_out/ncu_launches.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 6 -c 16 --csv --log-file gpurun_out/launches.csv python bench.py --steps 4 --warmup 3 --cpu-seconds 0.1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:stft2048 -s 4 -c 1 -o gpurun_out/prof_melfused -f python bench.py --steps 2 --warmup 3 --cpu-seconds 0.1 > gpurun_out/ncu_melfused.log 2>&1
ls gpurun_out
