#!/bin/bash
# end-of-round evidence: every GPU test (not the 2^32 sweep), smoke, entry-point timings, default bench line, launch list with DRAM
# bytes of the dominant kernel (stamped with the source fingerprint by scripts/summarize_profiles.py)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m "gpu and not slow" 2>&1 | tail -4
python __graft_entry__.py --smoke 2>&1 | tail -1
python scripts/gpu_time_ops.py > gpurun_out/time_ops.txt 2>&1; cat gpurun_out/time_ops.txt
python bench.py --steps 200 --warmup 5 --cpu-seconds 10 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; tail -2 gpurun_out/bench_cfg2.err; cut -c1-300 gpurun_out/bench_cfg2.json
ncu --cache-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 8 -c 6 --csv --log-file gpurun_out/launches_warm.csv python bench.py --steps 4 --warmup 3 --cpu-seconds 0.1 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --cpu-seconds 0.2 > gpurun_out/ncu_launches.log 2>&1
grep -c stft2048_pair gpurun_out/launches_warm.csv gpurun_out/launches.csv
