set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
timeout 300 python scratch/asm_bench.py 216,216,216 2>&1 | tail -8
timeout 600 python bench.py --steps 4 --warmup 3 > gpurun_out/r8_bench_10m.json 2> gpurun_out/r8_bench_10m.err; tail -3 gpurun_out/r8_bench_10m.err; cut -c1-1500 gpurun_out/r8_bench_10m.json
JB_RB_IDENTITY=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r8_bench_10m_noident.json 2> gpurun_out/r8_noident.err; tail -3 gpurun_out/r8_noident.err
timeout 300 python bench.py --dims 108,108,108 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r8_bench_1p26m.json 2> gpurun_out/r8_b.err; tail -3 gpurun_out/r8_b.err
JB_CHUNK_ROWS=256 timeout 300 python bench.py --dims 108,108,108 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r8_bench_1p26m_chunk256.json 2> gpurun_out/r8_c.err; tail -3 gpurun_out/r8_c.err
python - <<'PY'
import json
for f in ["r8_bench_10m", "r8_bench_10m_noident", "r8_bench_1p26m", "r8_bench_1p26m_chunk256"]:
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, "value", round(d["value"], 3), "e2e", round(d["e2e"]["value"], 3), "lin", d["linear_iterations"], "conv", d["converged"])
        for k, v in d["kernels"].items():
            print("   ", k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items()})
    except Exception as e:
        print(f, "FAILED", e)
PY
