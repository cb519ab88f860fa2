mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "refactorisation or ilu0 or bicgstab or newton_step or reordered" 2>&1 | tail -3
show() { python - "$1" <<'PY'
import json,sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); k = d["kernels"]
print(sys.argv[1], "value %.3f e2e %.3f | ilu_factor ms %.3f frac %.3f | assembly frac %.3f | fused frac %.3f" % (d["value"], d["e2e"]["value"], k["ilu_factor"]["ms_per_launch"], k["ilu_factor"]["frac"], k["assembly"]["frac"], k["bicgstab_iteration"]["frac"]))
PY
}
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r3d_bench_10m_1gpu.json 2> gpurun_out/r3d.err; tail -2 gpurun_out/r3d.err; show gpurun_out/r3d_bench_10m_1gpu.json
timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --dims 108,108,108 > gpurun_out/r3d_bench_1p26m.json 2> gpurun_out/r3d.err; show gpurun_out/r3d_bench_1p26m.json
