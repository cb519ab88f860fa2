# First GPU call of the next round (1 GPU, ~6 min): sanitizer passes over the TMA-staged kernels, then the regular suite.
# Usage: gpurun --timeout 900 -- 'bash scratch/next_round_first.sh'
set -x
mkdir -p gpurun_out
K="assembly_kernel_variants or spmv_matches or ilu0_factor_and_apply or bicgstab_matches or operator_identity or two_colour_stream"
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$K" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -5 gpurun_out/sanitizer_memcheck.log
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "assembly_kernel_variants or spmv_matches" > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -5 gpurun_out/sanitizer_racecheck.log
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -5
timeout 600 python bench.py --steps 4 --warmup 3 > gpurun_out/next_bench_10m.json 2> gpurun_out/next_bench_10m.err; python scratch/show.py gpurun_out/next_bench_10m.json
