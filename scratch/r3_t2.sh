set -x
mkdir -p gpurun_out
timeout 300 scratch/bb/barrier_bench > gpurun_out/r3_barrier_bench.txt 2>&1; cat gpurun_out/r3_barrier_bench.txt
timeout 600 python -m pytest tests/test_gpu_widen2.py -m gpu -q 2>&1 | tail -15
