mkdir -p gpurun_out
timeout 600 python bench.py --steps 4 --warmup 3 > gpurun_out/r02_bench_10m_1gpu.json 2> gpurun_out/r02_bench_10m_1gpu.err; tail -3 gpurun_out/r02_bench_10m_1gpu.err; python scratch/show.py gpurun_out/r02_bench_10m_1gpu.json
timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --dims 108,108,108 > gpurun_out/r02_bench_1p26m_1gpu.json 2> gpurun_out/r02_bench_1p26m.err; python scratch/show.py gpurun_out/r02_bench_1p26m_1gpu.json
python - <<'PY'
import json
for f in ("gpurun_out/r02_bench_10m_1gpu.json", "gpurun_out/r02_bench_1p26m_1gpu.json"):
    d = json.load(open(f)); k = d["kernels"].get("bicgstab_iteration")
    if k:
        print(f, "fused: ms/launch %.4f frac %.3f bytes %d grid %d" % (k["ms_per_launch"], k["frac"], k["algorithmic_bytes_per_launch"], k["grid_ctas"]))
        for n, p in k["phases"].items(): print("    %-14s %s" % (n, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in p.items()}))
PY
D=216,216,216
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_10m.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_launch.log 2>&1; tail -1 gpurun_out/ncu_launch.log | cut -c1-300
for k in bicgstab_rb_iteration_kernel twophase_assemble_pair_kernel ilu_factor_rb_kernel; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 12 -c 1 -f -o gpurun_out/r02_full10m_$k python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_$k.log 2>&1; tail -2 gpurun_out/ncu_$k.log | cut -c1-200
done
ls -la gpurun_out | grep r02_full
