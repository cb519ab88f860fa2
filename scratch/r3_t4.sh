mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --dims 108,108,108 > gpurun_out/r3_bench_1p26m_1gpu.json 2> gpurun_out/r3_bench_1p26m.err; tail -2 gpurun_out/r3_bench_1p26m.err
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r3_bench_10m_1gpu.json 2> gpurun_out/r3_bench_10m_1gpu.err; tail -2 gpurun_out/r3_bench_10m_1gpu.err
python - <<'PY'
import json
for f in ("gpurun_out/r3_bench_10m_1gpu.json", "gpurun_out/r3_bench_1p26m_1gpu.json"):
    d = json.loads(open(f).read().strip().splitlines()[-1]); k = d["kernels"].get("bicgstab_iteration")
    print(f, "value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "lin its", d["linear_iterations_per_newton"])
    if k:
        print("  fused: ms/launch %.4f frac %.3f bytes %d grid %d" % (k["ms_per_launch"], k["frac"], k["algorithmic_bytes_per_launch"], k["grid_ctas"]))
        for n, p in k["phases"].items(): print("    %-14s %s" % (n, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in p.items()}))
PY
