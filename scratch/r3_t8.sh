mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_widen2.py -m gpu -q -x 2>&1 | tail -4
timeout 600 python profiles/scripts/r02_widen_bench.py > gpurun_out/r02_widen_bench_10m.jsonl 2> gpurun_out/r02_widen_bench.err; tail -3 gpurun_out/r02_widen_bench.err; cat gpurun_out/r02_widen_bench_10m.jsonl
