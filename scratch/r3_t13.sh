mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "fused or bicgstab or newton or config3 or operator_identity or perform_step" 2>&1 | tail -3
show() { python - "$1" <<'PY'
import json,sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); k = d["kernels"]; f = k["bicgstab_iteration"]
print(sys.argv[1], "value %.3f e2e %.3f lin/newton %s | fused ms %.4f frac %.3f" % (d["value"], d["e2e"]["value"], d["linear_iterations_per_newton"], f["ms_per_launch"], f["frac"]))
print("   ", {n: round(p["ms"], 4) for n, p in f["phases"].items()})
PY
}
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r3f_bench_10m_1gpu.json 2> gpurun_out/r3f.err; tail -2 gpurun_out/r3f.err; show gpurun_out/r3f_bench_10m_1gpu.json
timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --dims 108,108,108 > gpurun_out/r3f_bench_1p26m.json 2> gpurun_out/r3f.err; show gpurun_out/r3f_bench_1p26m.json
