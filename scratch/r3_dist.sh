mkdir -p gpurun_out
NP=${NP:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$NP --master-addr 127.0.0.1 --master-port 29631 tests/dist_gpu_check.py > gpurun_out/r3_dist_raw_${NP}.log 2>&1; echo "dist_gpu_check rc=$?" | tee gpurun_out/r02b_dist_check_${NP}gpu.log
grep "dist_gpu_check\|rror\|Traceback\|FAILED" gpurun_out/r3_dist_raw_${NP}.log | tail -20 | tee -a gpurun_out/r02b_dist_check_${NP}gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$NP --master-addr 127.0.0.1 --master-port 29632 tests/dist_gpu_check.py 60,60,60 > gpurun_out/r3_dist_raw2_${NP}.log 2>&1; echo "dist_gpu_check 60^3 rc=$?" | tee -a gpurun_out/r02b_dist_check_${NP}gpu.log
grep "dist_gpu_check\|rror\|Traceback\|FAILED" gpurun_out/r3_dist_raw2_${NP}.log | tail -20 | tee -a gpurun_out/r02b_dist_check_${NP}gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$NP --master-addr 127.0.0.1 --master-port 29633 bench.py --gpus $NP --steps 4 --warmup 3 > gpurun_out/r02b_bench_10m_${NP}gpu.json 2> gpurun_out/r02b_bench_${NP}gpu.err; tail -3 gpurun_out/r02b_bench_${NP}gpu.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r02b_bench_10m_${NP}gpu.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "lin/newton", d.get("linear_iterations_per_newton"))
k = d["kernels"]
for n, v in k.items():
    if n != "bicgstab_iteration": print("  ", n, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items()})
f = k.get("bicgstab_iteration")
if f:
    print("  fused ms/launch %.4f frac %.3f" % (f["ms_per_launch"], f["frac"]))
    for n, p in f["phases"].items(): print("     %-14s %s" % (n, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in p.items()}))
PY
