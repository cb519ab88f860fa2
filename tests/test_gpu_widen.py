"""GPU parity of the pieces either side of the hot path: equations evaluated on the host and filled
into the device system through their GenericAutoDiffCache. Runs after test_gpu_parity.py."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dims", [(37, 29), (2, 3)])
def test_generic_cache_fill_matches_dedicated_heat_kernel(J, O, ctx, dims):
    """Host-evaluated entries of SimpleHeatEquation scattered by jb_generic_fill == the dedicated device assembly == oracle;
    the device BiCGStab then solves the system filled that way."""
    nx, ny = dims
    sim = J.HeatSimulator(ctx, nx, ny, 3.0, 2.0, rtol=1e-10)
    rng = np.random.default_rng(9)
    T = rng.uniform(0, 100, nx * ny); T0 = rng.uniform(0, 100, nx * ny)
    sim.T.set(T); sim.T0.set(T0)
    sim.assemble(0.25)
    nz_dev, r_dev = sim.jac.nonzeros(), sim.r.get()
    rowptr, colidx = sim.jac.pattern()
    vpos, dpos, pos, ent = O.generic_cache_heat(nx, ny, 3.0 / nx, 2.0 / ny, 0.25, T, T0, rowptr, colidx)
    fill = J.GenericAutoDiffCacheFill(sim.jac, 1, 1, vpos, pos, dpos)
    sim.jac.set_nonzeros(np.full_like(nz_dev, np.nan)); sim.r.set(np.full_like(r_dev, np.nan))
    fill.fill(ent, sim.r)
    nz_o = np.full_like(nz_dev, np.nan); r_o = np.full_like(r_dev, np.nan)
    O.generic_fill(1, 1, vpos, dpos, pos, ent, nz_o, r_o)
    assert np.array_equal(sim.jac.nonzeros(), nz_o) and np.array_equal(sim.r.get(), r_o)      # a scatter: bit-exact
    assert np.allclose(nz_o, nz_dev, rtol=1e-13) and np.allclose(r_o, r_dev, rtol=1e-12, atol=1e-12)
    ok, its, hist, st = J.linear_solve(sim.krylov, sim.r, sim.dx)
    assert ok and hist[-1] <= 1e-9 * hist[0] + 1e-12


def test_generic_cache_fill_bad_tables(J, ctx):
    A = J.build_sparse_matrix(ctx, [1, 1, 2, 2], [1, 2, 1, 2], 2, 1)
    with pytest.raises(J.JutulB200Error):
        J.GenericAutoDiffCacheFill(A, 1, 1, [1, 2, 3], [1, 99])            # position outside nzval
    with pytest.raises(J.JutulB200Error):
        J.GenericAutoDiffCacheFill(A, 1, 1, [1, 2, 3], [1, 4], [2, 2])     # diagonal slot outside the entity's slots
    f = J.GenericAutoDiffCacheFill(A, 1, 1, [1, 3, 5], [1, 2, 3, 4], [1, 4])
    r = ctx.zeros(2)
    f.fill(np.array([[10.0, 1.0], [10.0, 2.0], [20.0, 3.0], [20.0, 4.0]]), r)
    assert np.array_equal(A.nonzeros(), [1.0, 2.0, 3.0, 4.0]) and np.array_equal(r.get(), [10.0, 20.0])


def test_value_index_range_is_guarded(J, ctx):
    """Kernels index values with int32: nnz_blocks*bs*bs at or above 2^31 is refused when the matrix is created (the count
    is checked before the index arrays are read, so tiny arrays suffice here), not discovered as a wild write later."""
    import ctypes as C
    I = np.ones(4, dtype=np.int64)
    pi = I.ctypes.data_as(C.POINTER(C.c_int64))
    h = C.c_void_p()
    assert ctx.lib.jb_csr_create_from_coo(ctx.h, pi, pi, 2**29, 10, 2, C.byref(h)) == -4    # JB_ERR_UNSUPPORTED
    assert ctx.lib.jb_csr_create_from_coo(ctx.h, pi, pi, 4, 2**30, 2, C.byref(h)) == -4     # n*bs
    assert b"overflow" in ctx.lib.jb_last_error(ctx.h)


@pytest.mark.parametrize("solver,side,restart,memory", [("gmres", "right", True, 40), ("gmres", "left", True, 40), ("gmres", "none", True, 40),
                                                        ("fgmres", "right", False, 20), ("fgmres", "right", True, 40)])
def test_gmres_restart_and_fgmres_match_oracle(J, O, ctx, solver, side, restart, memory):
    """Restarted GMRES (the distributed call of the reference: restart = true, left preconditioning,
    ext/JutulPartitionedArraysExt/krylov.jl:67-74) and Krylov.jl fgmres!, against the oracle with the same options.
    Restarted GMRES converges slowly on this system (113-128 iterations at memory 40 for rtol 1e-6, none without a
    preconditioner), so after the first restarts only the counts and the final residual are compared."""
    from conftest import to_scipy
    from test_gpu_parity import _jacobian_on_gpu
    w, s, sim, nz, r = _jacobian_on_gpu(J, O, ctx, dims=(10, 9, 7))
    n = w["nc"]
    sim.jac.scale(sim.r, "diagonal")
    nz = sim.jac.nonzeros(); r = sim.r.get()
    prec = None if side == "none" else sim.prec
    itmax = 100 if side == "none" else 400
    kry = J.GenericKrylov(sim.jac, solver, prec, relative_tolerance=1e-6, precond_side="left" if side == "left" else "right",
                          max_iterations=itmax, memory=memory, restart=restart)
    ok, its, hist, st = J.linear_solve(kry, sim.r, sim.dx)
    ilu = O.ILU0(n, 2, s["rowptr"], s["colidx"]); ilu.factor(nz)
    x, st_o, its_o, hist_o = O.gmres(n, 2, s["rowptr"], s["colidx"], nz, r, ilu if side != "none" else None, side=side, rtol=1e-6, itmax=itmax,
                                     memory=memory, restart=restart, flexible=(solver == "fgmres"))
    m = min(len(hist), len(hist_o), memory + 5)
    assert np.allclose(hist[:m], hist_o[:m], rtol=1e-4, atol=1e-12 * hist[0])      # first pass and the first steps after the restart
    if side == "none":
        assert (not ok) and st == J.JB_NOT_CONVERGED and st_o == 1 and its == its_o == itmax
        assert np.allclose(hist, hist_o, rtol=1e-3)
        return
    assert ok and st == 0 and st_o == 0
    if restart:
        assert its > memory                    # at least one restart happened
    assert abs(its - its_o) <= max(2, (15 * its_o) // 100)
    dx = sim.dx.get()
    assert np.linalg.norm(dx + x) <= 1e-3 * np.linalg.norm(x)
    A = to_scipy(n, 2, s["rowptr"], s["colidx"], nz)
    assert np.linalg.norm(r + A @ dx) <= 1e-5 * np.linalg.norm(r)


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
    assert np.all(np.diff(hist) <= 1e-12 * hist[0])            # GMRES residual estimates never increase
    if kind == "spai0":
        return       # SPAI(0) scales rows by 1/|row|^2: too weak to reach 1e-6 in 300 iterations on this system (measured)
    A = to_scipy(n, 2, s["rowptr"], s["colidx"], nz)
    assert ok and np.linalg.norm(r + A @ sim.dx.get()) <= 1e-5 * np.linalg.norm(r)
    kb = J.GenericKrylov(sim.jac, "bicgstab", prec, relative_tolerance=1e-6, max_iterations=600)
    ok, its, hist, st = J.linear_solve(kb, sim.r, sim.dx)
    assert np.linalg.norm(r + A @ sim.dx.get()) <= 1e-4 * np.linalg.norm(r) or not ok
