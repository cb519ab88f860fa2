set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -6
D=216,216,216
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r14_launches_10m.csv python scratch/prof_kernels.py $D > gpurun_out/r14_ncu_launch.log 2>&1; tail -1 gpurun_out/r14_ncu_launch.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"twophase_assemble_pair|spmv_s2|ilu_sweep_s2|ilu_factor_rb|ilu_gather|bicg_update" -s 14 -c 12 -f -o gpurun_out/r14_full10m python scratch/prof_kernels.py $D > gpurun_out/r14_ncu_full.log 2>&1; tail -2 gpurun_out/r14_ncu_full.log
ncu -i gpurun_out/r14_full10m.ncu-rep --page raw --csv > gpurun_out/r14_full10m.raw.csv 2>/dev/null
timeout 900 python bench.py > gpurun_out/r14_bench_10m.json 2> gpurun_out/r14_bench_10m.err; tail -3 gpurun_out/r14_bench_10m.err
python scratch/show.py gpurun_out/r14_bench_10m.json
