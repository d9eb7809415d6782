#!/bin/bash
# ncu evidence for the tcgen05 pass-2 variant next to the default pair kernel (config 2 and config 3), FFMA2 latency
set -x
mkdir -p gpurun_out
scripts/micro/ffma2_latency > gpurun_out/ffma2_latency.txt 2>&1; cat gpurun_out/ffma2_latency.txt
TAC_MEL_VARIANT=2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:stft2048_pair -s 3 -c 1 -o gpurun_out/prof_pair_tc_cfg2 -f python scripts/gpu_mel_once.py cfg2 5 > gpurun_out/ncu_pair_tc.log 2>&1
TAC_MEL_VARIANT=2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:stft2048_pair -s 2 -c 1 -o gpurun_out/prof_pair_tc_cfg3 -f python scripts/gpu_mel_once.py cfg3 4 >> gpurun_out/ncu_pair_tc.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stft2048_pair -s 3 -c 1 -o gpurun_out/prof_pair_r02_cfg2 -f python scripts/gpu_mel_once.py cfg2 5 > gpurun_out/ncu_pair.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stft2048_pair -s 2 -c 1 -o gpurun_out/prof_pair_r02_cfg3 -f python scripts/gpu_mel_once.py cfg3 4 >> gpurun_out/ncu_pair.log 2>&1
tail -3 gpurun_out/ncu_pair_tc.log gpurun_out/ncu_pair.log
ls -la gpurun_out/*.ncu-rep
