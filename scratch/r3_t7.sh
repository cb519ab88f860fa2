mkdir -p gpurun_out
timeout 900 python profiles/scripts/r02_widen_bench.py > gpurun_out/r02_widen_bench_10m.jsonl 2> gpurun_out/r02_widen_bench.err; tail -3 gpurun_out/r02_widen_bench.err; cat gpurun_out/r02_widen_bench_10m.jsonl
