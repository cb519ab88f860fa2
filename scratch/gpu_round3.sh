python -m pytest tests -m gpu -q 2>&1 | tail -8
python bench.py --dims 100,100,100 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_1m.json 2> gpurun_out/bench_1m.err; tail -3 gpurun_out/bench_1m.err
python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_10m.json 2> gpurun_out/bench_10m.err; tail -3 gpurun_out/bench_10m.err
python bench.py --steps 3 --warmup 2 --no-cpu-baseline --ordering given > gpurun_out/bench_10m_given.json 2> gpurun_out/bench_10m_given.err; tail -3 gpurun_out/bench_10m_given.err
