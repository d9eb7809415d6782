set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:stft2048 -s 4 -c 1 -o gpurun_out/prof_melfused -f python bench.py --steps 2 --warmup 3 --cpu-seconds 0.1 > gpurun_out/ncu_melfused.log 2>&1
python scripts/gpu_fused_layouts.py
