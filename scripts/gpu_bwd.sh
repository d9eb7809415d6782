set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_backward.py -q -m gpu 2>&1 | tail -40 > gpurun_out/t_bwd.log; cat gpurun_out/t_bwd.log
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m "gpu and not slow" -x 2>&1 | tail -5
