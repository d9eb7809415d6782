# fused mel kernel: parity, A/B timing, launch list, one full capture
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m "gpu and not slow" -x -k "mel or fused or host_pipeline or stft_2048" 2>&1 | tail -15 > gpurun_out/t_fused.log; cat gpurun_out/t_fused.log
python scripts/gpu_fused_layouts.py > gpurun_out/fused_layouts.txt 2>&1; cat gpurun_out/fused_layouts.txt
python bench.py --steps 500 --warmup 5 --cpu-seconds 2 > gpurun_out/bench_fused.json 2> gpurun_out/bench_fused.err; cut -c1-2600 gpurun_out/bench_fused.json; tail -3 gpurun_out/bench_fused.err
TAC_MELSPEC_FUSED=0 python bench.py --steps 500 --warmup 5 --cpu-seconds 1 > gpurun_out/bench_pair.json 2>/dev/null; cut -c1-900 gpurun_out/bench_pair.json
python bench.py --steps 20 --warmup 3 --workload cfg3 --cpu-seconds 1 > gpurun_out/bench_cfg3_fused.json 2>/dev/null; cut -c1-700 gpurun_out/bench_cfg3_fused.json
ncu --cache-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -k regex:stft2048 -s 6 -c 8 --csv --log-file gpurun_out/launches_fused_warm.csv python bench.py --steps 6 --warmup 3 --cpu-seconds 0.1 > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stft2048 -s 4 -c 1 -o gpurun_out/prof_melfused -f python bench.py --steps 2 --warmup 3 --cpu-seconds 0.1 > gpurun_out/ncu_melfused.log 2>&1
ls -la gpurun_out
