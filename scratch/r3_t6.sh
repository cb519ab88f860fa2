set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_widen2.py -m gpu -q -x -k "fused_bicgstab" 2>&1 | tail -5
K="schur or adjoint or nfvm_law or tables_match or secondary_variable_graph_matches or fused_bicgstab"
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_widen2.py -m gpu -q -x -k "$K" > gpurun_out/r02b_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r02b_sanitizer_memcheck.log
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_widen2.py -m gpu -q -x -k "schur_operator or nfvm_law or secondary_variable_graph_matches or tables_match" > gpurun_out/r02b_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r02b_sanitizer_racecheck.log
timeout 600 python bench.py --config c2 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_bench_c2.json 2> gpurun_out/r02b_bench_c2.err; tail -2 gpurun_out/r02b_bench_c2.err; cut -c1-600 gpurun_out/r02b_bench_c2.json
timeout 600 python bench.py --config c3 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_bench_c3.json 2> gpurun_out/r02b_bench_c3.err; tail -2 gpurun_out/r02b_bench_c3.err; cut -c1-600 gpurun_out/r02b_bench_c3.json
