#!/bin/bash
mkdir -p gpurun_out
TAC_B200_LIB=$PWD/torchaudio_contrib_b200/lib/variants/libtac_w12.so timeout 300 ncu --set full --clock-control none --import-source on -k regex:stft2048_pair -s 2 -c 1 -o gpurun_out/prof_pair_w12_cfg3 -f python scripts/gpu_mel_once.py cfg3 4 > gpurun_out/ncu_w12.log 2>&1
tail -2 gpurun_out/ncu_w12.log
ls -la gpurun_out/prof_pair_w12_cfg3.ncu-rep
