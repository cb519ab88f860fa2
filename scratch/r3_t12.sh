mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
show() { python - "$1" <<'PY'
import json,sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); k = d["kernels"]
print(sys.argv[1], "value %.3f e2e %.3f | ilu_factor ms %.3f frac %.3f | assembly frac %.3f | fused frac %.3f" % (d["value"], d["e2e"]["value"], k["ilu_factor"]["ms_per_launch"], k["ilu_factor"]["frac"], k["assembly"]["frac"], k["bicgstab_iteration"]["frac"]))
PY
}
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r3e_bench_10m_1gpu.json 2> gpurun_out/r3e.err; tail -2 gpurun_out/r3e.err; show gpurun_out/r3e_bench_10m_1gpu.json
timeout 600 python profiles/scripts/r02_widen_bench.py > gpurun_out/r02_widen_bench_10m.jsonl 2> gpurun_out/r02_widen_bench.err; tail -3 gpurun_out/r02_widen_bench.err; head -2 gpurun_out/r02_widen_bench_10m.jsonl | cut -c1-200
