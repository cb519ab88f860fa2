set -x
python -m pytest tests -m gpu -q 2>&1 | tail -4
python bench.py --steps 4 --warmup 3 > gpurun_out/bench_10m.json 2> gpurun_out/bench_10m.err; tail -3 gpurun_out/bench_10m.err; cat gpurun_out/bench_10m.json | cut -c1-6000
python bench.py --impl reference --steps 1 > gpurun_out/bench_ref_10m.json 2> gpurun_out/bench_ref.err; tail -3 gpurun_out/bench_ref.err; cut -c1-1500 gpurun_out/bench_ref_10m.json
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_1m.csv python scratch/prof_kernels.py 100,100,100 > gpurun_out/ncu_launch.log 2>&1; tail -2 gpurun_out/ncu_launch.log
for k in twophase_assemble_kernel spmv_kernel ilu_forward_level ilu_backward_level ilu_factor_level; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/full_$k python scratch/prof_kernels.py 100,100,100 > gpurun_out/ncu_$k.log 2>&1; tail -1 gpurun_out/ncu_$k.log
done
ls -la gpurun_out
