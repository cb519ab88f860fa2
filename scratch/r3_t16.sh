mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_widen2.py -m gpu -q -x -k "poisson_adjoint" 2>&1 | tail -25
