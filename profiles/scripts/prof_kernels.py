"""Runs each hot kernel a few times on the given workload (multicolour device numbering) — target command for ncu captures."""
import sys
sys.path.insert(0, '.')
import numpy as np
import __graft_entry__ as g
J = g.load_package()
dims = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "100,100,100").split(",")]
w = J.workloads.unstructured_hex(*dims)
n = w["nc"]
ctx = J.B200Context(0)
sim = J.TwoPhaseSimulator(ctx, w["N"], n, w["Tf"], w["gdz"], w["pv"], w["params"], ordering="multicolor")
sim.set_forces(w["src_cells"], w["src_vals"])
sim.set_state(w["p0"], w["sw0"])
x = ctx.transfer(np.random.default_rng(0).standard_normal(2 * n)); y = ctx.zeros(2 * n)
for rep in range(3):
    sim.law.update_equation_and_linearized_system(sim.p, sim.s, sim.M0, w["dt"], sim.r)
    sim.jac.mul(y, x)
    sim.prec.update_preconditioner()
    sim.prec.apply(y, x)
ok, its, hist, st = J.linear_solve(sim.krylov, sim.r, sim.dx)
print("done", its)
