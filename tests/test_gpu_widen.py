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


@pytest.mark.parametrize("layout", ["equation_major", "entity_major"])
def test_scalar_jacobian_layouts_on_device(J, O, ctx, layout):
    """a3: the reference's scalar layouts. The two-phase Jacobian as a scalar CSR of 2 nc rows, dofs numbered by
    EquationMajorLayout (index = n (e-1) + c) or EntityMajorLayout (index = 2 (c-1) + e) — the oracle's index formulas,
    pinned to test/adjoints/utils.jl:57-68 by test_layout_goldens. Pattern bit-exact; SpMV, ILU(0) and the Krylov solve on the
    bs = 1 device path reproduce the 2x2-block results up to the dof permutation."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    from test_gpu_parity import _jacobian_on_gpu
    w, s, sim, nz, r = _jacobian_on_gpu(J, O, ctx, dims=(7, 6, 5))
    n = w["nc"]
    nz = sim.jac.nonzeros(); r = sim.r.get()
    rp, ci = s["rowptr"], s["colidx"]
    idx = (lambda c, d: O.index_equation_major(c, d, n)) if layout == "equation_major" else (lambda c, d: O.index_entity_major(c, d, 2))
    rows = np.repeat(np.arange(n), np.diff(rp)); cols = ci - 1
    I, Jc, V = [], [], []
    for e in range(2):
        for q in range(2):
            I.append(np.array([idx(c + 1, e + 1) for c in rows])); Jc.append(np.array([idx(c + 1, q + 1) for c in cols]))
            V.append(nz[2 * q + e::4])                       # block entry (eq e, partial q), column-major blocks
    I, Jc, V = np.concatenate(I).astype(np.int64), np.concatenate(Jc).astype(np.int64), np.concatenate(V)
    M = sp.coo_matrix((V, (I - 1, Jc - 1)), shape=(2 * n, 2 * n)).tocsr(); M.sort_indices()
    A = J.build_sparse_matrix(ctx, I, Jc, 2 * n, 1)
    rps, cis = A.pattern()
    rpo, cio = O.csr_from_coo(I, Jc, 2 * n)
    assert np.array_equal(rps, rpo) and np.array_equal(cis, cio) and np.array_equal(rps, M.indptr + 1) and np.array_equal(cis, M.indices + 1)
    A.set_nonzeros(M.data)
    P = np.array([idx(c + 1, e + 1) - 1 for c in range(n) for e in range(2)])     # block dof 2c+e -> layout dof
    rng = np.random.default_rng(5)
    x = rng.standard_normal(2 * n); xs = np.zeros(2 * n); xs[P] = x
    y = ctx.zeros(2 * n)
    A.mul(y, ctx.transfer(xs))
    yb = O.spmv(n, 2, rp, ci, nz, x)
    assert np.abs(y.get()[P] - yb).max() <= 1e-13 * np.abs(yb).max()
    prec = J.ILUZeroPreconditioner(A); assert prec.update_preconditioner() == 0
    ilu = O.ILU0(2 * n, 1, rps, cis); assert ilu.factor(M.data) == 0
    z = ctx.zeros(2 * n)
    prec.apply(z, ctx.transfer(xs))
    zo = ilu.solve(xs)
    assert np.abs(z.get() - zo).max() <= 1e-8 * np.abs(zo).max()
    rs = np.zeros(2 * n); rs[P] = r
    kry = J.GenericKrylov(A, "bicgstab", prec, relative_tolerance=1e-10, max_iterations=500)
    dx = ctx.zeros(2 * n)
    ok, its, hist, st = J.linear_solve(kry, ctx.transfer(rs), dx)
    assert ok and st == 0
    xb = spla.spsolve(to_scipy_block(n, rp, ci, nz), r)       # the block system's own solution
    assert np.linalg.norm(-dx.get()[P] - xb) <= 1e-5 * np.linalg.norm(xb)


def to_scipy_block(n, rp, ci, nz):
    from conftest import to_scipy
    return to_scipy(n, 2, rp, ci, nz).tocsc()
