#!/bin/bash
# first run of the tcgen05 pass-2 variant: agreement with the other kernels, timing (bounded: a hang must not eat the box)
mkdir -p gpurun_out
timeout 300 python scripts/gpu_pair_ab.py 2>&1 | tee gpurun_out/pair_ab_tc.txt | tail -40
