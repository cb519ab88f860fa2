mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none -k regex:ilu_factor_rb2_kernel -s 1 -c 1 -f -o gpurun_out/r3_rb2 python scratch/prof_kernels.py 216,216,216 > gpurun_out/r3_ncu_rb2.log 2>&1; tail -2 gpurun_out/r3_ncu_rb2.log
ncu -i gpurun_out/r3_rb2.ncu-rep --page raw --csv > gpurun_out/r3_rb2_raw.csv 2>/dev/null; python scratch/ncu_summary.py gpurun_out/r3_rb2_raw.csv
