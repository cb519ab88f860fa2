import sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import __graft_entry__ as g
J = g.load_package()
import oracle as O
from conftest import oracle_system
ctx = J.B200Context(0)
for dims, rtol in (((13, 11, 7), 1e-8), ((40, 40, 40), 1e-6), ((100,100,100), 1e-6)):
    w = J.workloads.unstructured_hex(*dims)
    n = w["nc"]
    t0=time.time()
    sim = J.TwoPhaseSimulator(ctx, w["N"], n, w["Tf"], w["gdz"], w["pv"], w["params"], rtol=rtol, max_linear_iterations=300)
    print(dims, "setup s", time.time()-t0, sim.prec.info())
    sim.set_forces(w["src_cells"], w["src_vals"])
    sim.set_state(w["p0"], w["sw0"])
    rng = np.random.default_rng(11)
    p = w["p0"] * (1 + 1e-3 * rng.standard_normal(n))
    sim.p.set(p)
    sim.law.update_equation_and_linearized_system(sim.p, sim.s, sim.M0, w["dt"], sim.r)
    t0=time.time()
    ok, its, hist, st = J.linear_solve(sim.krylov, sim.r, sim.dx)
    print("gpu", ok, its, st, "time", time.time()-t0, hist[:6], hist[-3:])
    if n <= 70000:
        s = oracle_system(O, w)
        nz = sim.jac.nonzeros(); r = sim.r.get()
        ilu = O.ILU0(n, 2, s["rowptr"], s["colidx"]); ilu.factor(nz)
        x, st_o, its_o, hist_o = O.bicgstab(n, 2, s["rowptr"], s["colidx"], nz, r, ilu, rtol=rtol, itmax=300)
        print("cpu", st_o, its_o, hist_o[:6], hist_o[-3:])
        print("dx diff", np.linalg.norm(sim.dx.get()+x)/np.linalg.norm(x))
