mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json,sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); k = d["kernels"].get("bicgstab_iteration")
print(sys.argv[1], "value %.3f ms/step %.2f e2e %.3f (ms/step %.2f) lin its %s fused ms/launch %.4f frac %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["linear_iterations_per_newton"], k["ms_per_launch"], k["frac"]))
PY
}
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r3b_bench_10m_1gpu.json 2> gpurun_out/r3b_bench_10m_1gpu.err; tail -2 gpurun_out/r3b_bench_10m_1gpu.err; show gpurun_out/r3b_bench_10m_1gpu.json
for c in 6 5 4 3 2; do
JB_PK_CTAS_PER_SM=$c timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --dims 108,108,108 > gpurun_out/r3b_bench_1p26m_ctas$c.json 2> gpurun_out/r3b_1p26m.err; tail -1 gpurun_out/r3b_1p26m.err; show gpurun_out/r3b_bench_1p26m_ctas$c.json
done
