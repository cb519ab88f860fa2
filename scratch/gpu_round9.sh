set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
D=216,216,216
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"twophase_assemble_tma|ilu_factor_rb|ilu_gather" -s 3 -c 3 -f -o gpurun_out/r9_full10m_asm python scratch/prof_kernels.py $D > gpurun_out/r9_ncu_asm.log 2>&1; tail -2 gpurun_out/r9_ncu_asm.log
timeout 600 ncu --set full --clock-control none -k regex:"spmv_stream_kernel|ilu_sweep_stream" -s 8 -c 4 -f -o gpurun_out/r9_full10m_spmv python scratch/prof_kernels.py $D > gpurun_out/r9_ncu_spmv.log 2>&1; tail -2 gpurun_out/r9_ncu_spmv.log
for f in r9_full10m_asm r9_full10m_spmv; do
  ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.raw.csv 2>/dev/null
done
ls -la gpurun_out/
