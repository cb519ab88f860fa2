python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -3
python -m pytest tests -m gpu -q 2>&1 | tail -3
(time python bench.py) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -4 gpurun_out/bench_default.err
(time python bench.py --impl reference) > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -4 gpurun_out/bench_reference.err
D=216,216,216
ncu --set full --clock-control none -k regex:"ilu_sweep_stream_kernel|ilu_iso_kernel" -s 4 -c 2 -f -o gpurun_out/final10m_ilu python scratch/prof_kernels.py $D > gpurun_out/ncu_ilu_final.log 2>&1; tail -1 gpurun_out/ncu_ilu_final.log
ncu --set full --clock-control none -k regex:twophase_assemble_stream -s 1 -c 1 -f -o gpurun_out/final10m_asm python scratch/prof_kernels.py $D > gpurun_out/ncu_asm_final.log 2>&1; tail -1 gpurun_out/ncu_asm_final.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_10m_final.csv python scratch/prof_kernels.py $D > gpurun_out/ncu_launch_final.log 2>&1; tail -1 gpurun_out/ncu_launch_final.log
