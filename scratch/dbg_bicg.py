import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import __graft_entry__ as g
J = g.load_package()
import oracle as O
O.build()
from conftest import oracle_system, to_scipy
import scipy.sparse.linalg as spla
ctx = J.B200Context(0)
from test_gpu_parity import _jacobian_on_gpu
w, s, sim, nz, r = _jacobian_on_gpu(J, O, ctx, dims=(13, 11, 7))
n = w["nc"]
nz = sim.jac.nonzeros(); r = sim.r.get()
A = to_scipy(n, 2, s["rowptr"], s["colidx"], nz)
xd = spla.spsolve(A.tocsc(), r)
ilu = O.ILU0(n, 2, s["rowptr"], s["colidx"]); ilu.factor(nz)
for rtol in (1e-3, 1e-8):
    for side in ("right", "left"):
        kry = J.GenericKrylov(sim.jac, "bicgstab", sim.prec, relative_tolerance=rtol, precond_side=side, max_iterations=200)
        ok, its, hist, st = J.linear_solve(kry, sim.r, sim.dx)
        x, st_o, its_o, hist_o = O.bicgstab(n, 2, s["rowptr"], s["colidx"], nz, r, ilu, side=side, rtol=rtol, itmax=200)
        dx = sim.dx.get()
        eg = np.linalg.norm(dx + xd) / np.linalg.norm(xd); eo = np.linalg.norm(x - xd) / np.linalg.norm(xd)
        rg = np.linalg.norm(r + A @ dx) / np.linalg.norm(r); ro = np.linalg.norm(r - A @ x) / np.linalg.norm(r)
        print(f"rtol {rtol:g} {side}: its gpu {its} oracle {its_o}  err gpu {eg:.3e} oracle {eo:.3e}  true relres gpu {rg:.3e} oracle {ro:.3e}  hist_end gpu {hist[-1]/hist[0]:.3e} oracle {hist_o[-1]/hist_o[0]:.3e}")
