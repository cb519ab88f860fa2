D=216,216,216
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_10m.csv python scratch/prof_kernels.py $D > gpurun_out/ncu_launch.log 2>&1; tail -1 gpurun_out/ncu_launch.log
for k in twophase_assemble_kernel spmv_stream_kernel ilu_sweep_stream_kernel ilu_factor_level_kernel bicg_update2_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 5 -c 2 -f -o gpurun_out/full10m_$k python scratch/prof_kernels.py $D > gpurun_out/ncu_$k.log 2>&1; tail -1 gpurun_out/ncu_$k.log
done
ls -la gpurun_out | grep full10m
