#!/bin/bash
# programmatic dependent launch of the pair kernel: parity / graph tests, then config 2 and 3 with it on and off
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_peers.py -q -x -k "pair or fused or mel or peer or full_size or graph or repeated" 2>&1 | tail -4
for v in 1 0; do echo -n "TAC_PAIR_PDL=$v  "; TAC_PAIR_PDL=$v python scripts/gpu_time_variant.py; done 2>&1 | tee gpurun_out/pdl_ab.txt
