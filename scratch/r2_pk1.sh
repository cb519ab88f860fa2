set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --fixed-state > gpurun_out/pk1_bench_10m_fixed.json 2> gpurun_out/pk1_bench_10m_fixed.err; tail -3 gpurun_out/pk1_bench_10m_fixed.err; python scratch/show.py gpurun_out/pk1_bench_10m_fixed.json
JB_PERSISTENT=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --fixed-state > gpurun_out/pk1_bench_10m_fixed_off.json 2> gpurun_out/pk1_bench_10m_fixed_off.err; python scratch/show.py gpurun_out/pk1_bench_10m_fixed_off.json
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/pk1_bench_10m_adv.json 2> gpurun_out/pk1_bench_10m_adv.err; tail -3 gpurun_out/pk1_bench_10m_adv.err; python scratch/show.py gpurun_out/pk1_bench_10m_adv.json
