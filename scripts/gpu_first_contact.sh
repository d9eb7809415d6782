set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "mulaw and not exhaustive" 2>&1 | tail -15 > gpurun_out/t_mulaw.log; cat gpurun_out/t_mulaw.log
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "complex_norm or amplitude_db or stft or cfg1 or spectrogram_db" 2>&1 | tail -30 > gpurun_out/t_stft.log; cat gpurun_out/t_stft.log
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "stages or apply_filterbank" 2>&1 | tail -30 > gpurun_out/t_fb.log; cat gpurun_out/t_fb.log
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "mel_ or meldb or full_size or host_pipeline or stretch" 2>&1 | tail -40 > gpurun_out/t_mel.log; cat gpurun_out/t_mel.log
