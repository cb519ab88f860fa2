"""GPU tests of code written after the GPU budget of round 1 was spent: compiled for sm_100a but not yet run on a
B200. They are skipped unless JB_RUN_UNVERIFIED=1, so a round-end `pytest -m gpu` reports them as skipped rather than
passing or failing on code nobody has executed. First thing to run in the next round:
    JB_RUN_UNVERIFIED=1 python -m pytest tests/test_gpu_unverified.py -m gpu -q"""
import os

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(os.environ.get("JB_RUN_UNVERIFIED") != "1", reason="not yet run on a B200 (JB_RUN_UNVERIFIED=1)")]


@pytest.mark.parametrize("kind", ["jacobi", "spai0"])
def test_diagonal_preconditioners_match_oracle(J, O, ctx, kind):
    """JacobiPreconditioner / SPAI0Preconditioner (src/linsolve/precond/jacobi.jl, spai.jl) behind the ILU handle:
    factors and apply! against the oracle, then as the preconditioner of the device GMRES."""
    from conftest import to_scipy
    from test_gpu_parity import _jacobian_on_gpu
    w, s, sim, nz, r = _jacobian_on_gpu(J, O, ctx, dims=(10, 9, 7))
    n = w["nc"]
    sim.jac.scale(sim.r, "diagonal")
    nz = sim.jac.nonzeros(); r = sim.r.get()
    prec = J.JacobiPreconditioner(sim.jac) if kind == "jacobi" else J.SPAI0Preconditioner(sim.jac)
    assert prec.update_preconditioner() == 0
    D_o = O.jacobi_factor(n, 2, s["rowptr"], s["colidx"], nz) if kind == "jacobi" else O.spai0_factor(n, 2, s["rowptr"], s["colidx"], nz)
    D_g = prec.factors()["D"]
    assert np.abs(D_g - D_o).max() <= 1e-13 * np.abs(D_o).max()
    y = np.random.default_rng(6).standard_normal(2 * n)
    x = ctx.zeros(2 * n)
    prec.apply(x, ctx.transfer(y))
    xo = O.diagonal_apply(D_o, y, 2)
    assert np.abs(x.get() - xo).max() <= 1e-13 * np.abs(xo).max()
    kry = J.GenericKrylov(sim.jac, "gmres", prec, relative_tolerance=1e-6, max_iterations=300)
    ok, its, hist, st = J.linear_solve(kry, sim.r, sim.dx)
    A = to_scipy(n, 2, s["rowptr"], s["colidx"], nz)
    assert ok and np.linalg.norm(r + A @ sim.dx.get()) <= 1e-5 * np.linalg.norm(r)
    kb = J.GenericKrylov(sim.jac, "bicgstab", prec, relative_tolerance=1e-6, max_iterations=600)
    ok, its, hist, st = J.linear_solve(kb, sim.r, sim.dx)
    assert np.linalg.norm(r + A @ sim.dx.get()) <= 1e-4 * np.linalg.norm(r) or not ok
