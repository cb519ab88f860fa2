set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -6
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r13_bench_10m.json 2> gpurun_out/r13_bench_10m.err; tail -3 gpurun_out/r13_bench_10m.err
timeout 300 python bench.py --dims 108,108,108 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r13_bench_1p26m.json 2> gpurun_out/r13_b.err; tail -3 gpurun_out/r13_b.err
python scratch/show.py gpurun_out/r13_bench_10m.json gpurun_out/r13_bench_1p26m.json
