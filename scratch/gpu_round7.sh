python -m pytest tests -m gpu -q 2>&1 | tail -6
python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_10m_b.json 2> gpurun_out/bench_10m_b.err; tail -3 gpurun_out/bench_10m_b.err
