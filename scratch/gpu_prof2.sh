D=216,216,216
ncu --set full --clock-control none --import-source on -k regex:twophase_assemble_stream -s 1 -c 1 -f -o gpurun_out/full10m_asm python scratch/prof_kernels.py $D > gpurun_out/ncu_asm.log 2>&1; tail -1 gpurun_out/ncu_asm.log
ncu --set full --clock-control none -k regex:"ilu_sweep_stream_kernel|ilu_light_level_kernel" -s 4 -c 4 -f -o gpurun_out/full10m_ilu4 python scratch/prof_kernels.py $D > gpurun_out/ncu_ilu4.log 2>&1; tail -1 gpurun_out/ncu_ilu4.log
