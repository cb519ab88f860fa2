set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -25
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"twophase_assemble_pair" -s 1 -c 1 -f -o gpurun_out/r16_full10m_asm python scratch/prof_kernels.py 216,216,216 > gpurun_out/r16_ncu.log 2>&1; tail -1 gpurun_out/r16_ncu.log
ncu -i gpurun_out/r16_full10m_asm.ncu-rep --page raw --csv > gpurun_out/r16_full10m_asm.raw.csv 2>/dev/null
