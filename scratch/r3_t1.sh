set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv
timeout 600 python -m pytest tests/test_gpu_widen2.py -m gpu -q 2>&1 | tail -60 > gpurun_out/r3_widen2.log; tail -40 gpurun_out/r3_widen2.log
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_widen2.py 2>&1 | tail -8
