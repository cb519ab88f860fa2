import os, sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import __graft_entry__ as g
J = g.load_package()
ctx = J.B200Context(0)
w = J.workloads.unstructured_hex(24, 20, 18)
res = {}
for pers in ("1", "0"):
    for ident in ("1", "0"):
        os.environ["JB_PERSISTENT"] = pers; os.environ["JB_RB_IDENTITY"] = ident
        sim = J.TwoPhaseSimulator(ctx, w["N"], w["nc"], w["Tf"], w["gdz"], w["pv"], w["params"], ordering="multicolor", rtol=1e-8, max_linear_iterations=300)
        sim.set_forces(w["src_cells"], w["src_vals"]); sim.set_state(w["p0"], w["sw0"])
        conv, err, rep = sim.perform_step(w["dt"])
        h = rep["linear_residuals"]
        res[(pers, ident)] = (rep["linear_iterations"], h, sim.dx.get())
        print("persistent", pers, "identity", ident, "its", rep["linear_iterations"], "status", rep["linear_status"], "info", sim.krylov.identity_info(), "rel", h[-1] / h[0])
        print("   hist", " ".join("%.6e" % x for x in h[:8]))
        print("   hist@", " ".join("%d:%.3e" % (k, h[k]) for k in range(10, len(h), 20)))
ref = res[("0", "0")]
for k, v in res.items():
    n = min(len(v[1]), len(ref[1]))
    rel = np.abs(v[1][:n] - ref[1][:n]) / ref[1][:n]
    first_bad = int(np.argmax(rel > 1e-6)) if (rel > 1e-6).any() else -1
    print(k, "dx diff vs old/noident", np.linalg.norm(v[2] - ref[2]) / np.linalg.norm(ref[2]), "history deviates >1e-6 at iteration", first_bad)
