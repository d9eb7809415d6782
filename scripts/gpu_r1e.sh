# round-1 evidence refresh: all GPU tests (not the 2^32 sweep), benches, launch lists, one full capture, backward timing
set -x
mkdir -p gpurun_out; rm -f gpurun_out/launches*.csv
timeout 1200 python -m pytest tests -q -m "gpu and not slow" 2>&1 | tail -6 > gpurun_out/t_all.log; cat gpurun_out/t_all.log
python __graft_entry__.py --smoke 2>&1 | tail -2
python scripts/gpu_time_backward.py > gpurun_out/time_backward.txt 2>&1; cat gpurun_out/time_backward.txt
python bench.py --steps 1000 --warmup 5 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; cut -c1-300 gpurun_out/bench_cfg2.json; tail -3 gpurun_out/bench_cfg2.err
python bench.py --steps 20 --warmup 3 --workload cfg3 --cpu-seconds 4 > gpurun_out/bench_cfg3.json 2>/dev/null; cut -c1-300 gpurun_out/bench_cfg3.json
python bench.py --workload mulaw --steps 20 > gpurun_out/bench_mulaw.json 2>/dev/null; cut -c1-300 gpurun_out/bench_mulaw.json
python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_reference.json 2>/dev/null; cut -c1-300 gpurun_out/bench_reference.json
python scripts/gpu_time_ops.py > gpurun_out/time_ops.txt 2>&1; tail -12 gpurun_out/time_ops.txt
ncu --cache-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -s 6 -c 16 --csv --log-file gpurun_out/launches_warm.csv python bench.py --steps 4 --warmup 3 --cpu-seconds 0.1 > gpurun_out/ncu_launches.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 6 -c 16 --csv --log-file gpurun_out/launches.csv python bench.py --steps 4 --warmup 3 --cpu-seconds 0.1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:stft2048_kernel -s 4 -c 1 -o gpurun_out/prof_melfused -f python bench.py --steps 2 --warmup 3 --cpu-seconds 0.1 > gpurun_out/ncu_melfused.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stft2048_backward -s 2 -c 1 -o gpurun_out/prof_melbwd -f python scripts/gpu_time_backward.py > gpurun_out/ncu_melbwd.log 2>&1
ls gpurun_out
