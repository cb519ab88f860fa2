"""Times the fused two-phase assembly kernel variants (JB_ASM_VARIANT 0 = TMA-staged, 1 = register stream,
2 = lane-per-half-face) on one workload with CUDA events on the library's stream; prints GB/s against the
algorithmic bytes. Also checks that the variants agree."""
import os
import sys
sys.path.insert(0, '.')
import numpy as np
import __graft_entry__ as g
J = g.load_package()
dims = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "100,100,100").split(",")]
w = J.workloads.unstructured_hex(*dims)
n, nf = w["nc"], w["nf"]
alg = J.workloads.algorithmic_bytes(n, nf, 2)["assembly"]
ctx = J.B200Context(0)
sim = J.TwoPhaseSimulator(ctx, w["N"], n, w["Tf"], w["gdz"], w["pv"], w["params"], ordering="multicolor")
sim.set_forces(w["src_cells"], w["src_vals"])
sim.set_state(w["p0"], w["sw0"])
lib = ctx.lib
ref = None
for v, ng in (("0", "36"), ("0", "27"), ("0", "37"), ("0", "4"), ("0", "36")):
    os.environ["JB_ASM_VARIANT"] = v
    os.environ["JB_ASM_NG"] = ng
    for _ in range(3):
        sim.law.update_equation_and_linearized_system(sim.p, sim.s, sim.M0, w["dt"], sim.r)
    reps = 10
    with J.DeviceProfile(ctx) as prof:
        for _ in range(reps):
            sim.law.update_equation_and_linearized_system(sim.p, sim.s, sim.M0, w["dt"], sim.r)
        cls = prof.collect()
    t, c = cls["assembly"]
    ts, cs = cls["state"]
    nz = sim.jac.nonzeros()
    if ref is None:
        ref = nz
    d = np.abs(nz - ref).max() / np.abs(ref).max()
    print(f"variant {v} ng {ng}: assembly {t / c * 1e3:8.1f} us/launch  {alg / (t / c * 1e-3) / 1e9:7.1f} GB/s  frac {alg / (t / c * 1e-3) / 1e9 / 6549.4:.3f}"
          f"   state {ts / cs * 1e3:6.1f} us   max rel diff vs variant 0: {d:.2e}", flush=True)
