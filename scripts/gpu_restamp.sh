#!/bin/bash
# DRAM bytes per launch of the dominant kernel (steady state of bench.py) -> gpurun_out/launches_warm.csv;
# scripts/summarize_profiles.py r02 then writes profiles/r02_melfused_dram_bytes.json stamped with the source fingerprint
mkdir -p gpurun_out
ncu --cache-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 8 -c 6 --csv --log-file gpurun_out/launches_warm.csv python bench.py --steps 4 --warmup 3 --cpu-seconds 0.1 --skip-extras > /dev/null 2>&1
grep -c stft2048_pair gpurun_out/launches_warm.csv
