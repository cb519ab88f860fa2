mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
show() { python - "$1" <<'PY'
import json,sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); k = d["kernels"]
print(sys.argv[1], "value %.3f e2e %.3f lin/newton %s | ilu_factor ms %.3f frac %.3f | assembly ms %.3f frac %.3f | fused frac %.3f" % (d["value"], d["e2e"]["value"], d["linear_iterations_per_newton"], k["ilu_factor"]["ms_per_launch"], k["ilu_factor"]["frac"], k["assembly"]["ms_per_launch"], k["assembly"]["frac"], k["bicgstab_iteration"]["frac"]))
PY
}
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r3c_bench_10m_1gpu.json 2> gpurun_out/r3c.err; tail -2 gpurun_out/r3c.err; show gpurun_out/r3c_bench_10m_1gpu.json
JB_ILU_FACTOR_RB1=1 timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r3c_bench_10m_1gpu_rb1.json 2> gpurun_out/r3c.err; tail -2 gpurun_out/r3c.err; show gpurun_out/r3c_bench_10m_1gpu_rb1.json
timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --dims 108,108,108 > gpurun_out/r3c_bench_1p26m.json 2> gpurun_out/r3c.err; show gpurun_out/r3c_bench_1p26m.json
