#!/bin/bash
# e2e of the config-2 step: one H2D stream against two alternating ones, a few slice sizes
for st in 1 2; do for mb in default 3 5; do
  echo -n "h2d_streams $st  "; TAC_HOST_H2D_STREAMS=$st python scripts/gpu_e2e_slices.py $mb
done; done 2>&1 | tee gpurun_out/e2e_streams.txt
