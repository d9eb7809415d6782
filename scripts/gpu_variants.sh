#!/bin/bash
mkdir -p gpurun_out
for so in "$@"; do TAC_B200_LIB=$PWD/torchaudio_contrib_b200/lib/variants/$so timeout 200 python scripts/gpu_time_variant.py; done 2>&1 | tee -a gpurun_out/variants_r02b.txt
timeout 200 python scripts/gpu_time_variant.py 2>&1 | tee -a gpurun_out/variants_r02b.txt
