"""Oracle self-checks where the reference holds no golden vector ("parity unpinned" items):
dense-algebra identities, scipy direct solves, finite differences. No GPU."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

from conftest import oracle_system, to_scipy


def _twophase_system(O, J, dims=(6, 5, 4), permute=True, perturb=True):
    w = J.workloads.unstructured_hex(*dims, permute=permute)
    s = oracle_system(O, w)
    M0 = O.mass_2ph(w["pv"], w["params"], w["p0"], w["sw0"])
    p = w["p0"] * (1 + (1e-3 * np.sin(np.arange(w["nc"])) if perturb else 0))
    nz, r = O.assemble_2ph(s["hf"], s["diag_pos"], s["hf_pos"], w["Tf"], w["gdz"], w["pv"], w["params"], p, w["sw0"], M0, w["dt"],
                           s["colidx"].shape[0], w["src_cells"], w["src_vals"])
    return w, s, M0, p, nz, r


def test_spmv_matches_scipy(O, J):
    w, s, M0, p, nz, r = _twophase_system(O, J)
    A = to_scipy(w["nc"], 2, s["rowptr"], s["colidx"], nz)
    x = np.random.default_rng(0).standard_normal(2 * w["nc"])
    y = O.spmv(w["nc"], 2, s["rowptr"], s["colidx"], nz, x)
    assert np.allclose(y, A @ x, rtol=1e-13, atol=1e-13 * np.abs(y).max())
    y2 = O.spmv(w["nc"], 2, s["rowptr"], s["colidx"], nz, x, alpha=2.0, beta=0.5, y=y.copy())
    assert np.allclose(y2, 2.0 * (A @ x) + 0.5 * y, rtol=1e-13, atol=1e-13 * np.abs(y).max())


def test_jacobian_finite_differences(O, J):
    """The AD Jacobian (local-perspective partials scattered as the reference does) equals dR/dx."""
    w, s, M0, p, nz, r = _twophase_system(O, J, dims=(4, 3, 3))
    A = to_scipy(w["nc"], 2, s["rowptr"], s["colidx"], nz).toarray()
    nc = w["nc"]
    rng = np.random.default_rng(3)
    for c in rng.choice(nc, 8, replace=False):
        for var, h in ((0, 10.0), (1, 1e-7)):
            pp, ss = p.copy(), w["sw0"].copy()
            pm, sm = p.copy(), w["sw0"].copy()
            if var == 0:
                pp[c] += h; pm[c] -= h
            else:
                ss[c] += h; sm[c] -= h
            rp = O.residual_2ph(s["hf"], w["Tf"], w["gdz"], w["pv"], w["params"], pp, ss, M0, w["dt"])
            rm = O.residual_2ph(s["hf"], w["Tf"], w["gdz"], w["pv"], w["params"], pm, sm, M0, w["dt"])
            fd = (rp - rm) / (2 * h)
            col = A[:, 2 * c + var]
            assert np.allclose(fd, col, rtol=2e-5, atol=1e-6 * np.abs(col).max())


def test_residual_identity(O, J):
    # test/test_systems/helper.jl:3-18: model_residual (no AD) == LinearizedSystem.r
    w, s, M0, p, nz, r = _twophase_system(O, J)
    r2 = O.residual_2ph(s["hf"], w["Tf"], w["gdz"], w["pv"], w["params"], p, w["sw0"], M0, w["dt"])
    src = np.zeros_like(r2)
    for c, v in zip(w["src_cells"], w["src_vals"]):
        src[2 * (c - 1):2 * c] += v
    assert np.allclose(r, r2 + src, rtol=1e-14, atol=1e-14 * np.abs(r).max())


@pytest.mark.parametrize("bs", [1, 2, 3])
def test_ilu0_exact_on_tridiagonal(O, bs):
    """ILU(0) of a (block) tridiagonal matrix has no dropped fill: L*U == A and the solve is exact."""
    n = 40
    rng = np.random.default_rng(bs)
    I = np.concatenate([np.arange(1, n + 1), np.arange(2, n + 1), np.arange(1, n)])
    Jc = np.concatenate([np.arange(1, n + 1), np.arange(1, n), np.arange(2, n + 1)])
    rowptr, colidx = O.csr_from_coo(I, Jc, n)
    nz = rng.standard_normal(colidx.shape[0] * bs * bs) * 0.3
    for row in range(n):   # diagonally dominant blocks
        for k in range(rowptr[row] - 1, rowptr[row + 1] - 1):
            if colidx[k] - 1 == row:
                blk = nz[k * bs * bs:(k + 1) * bs * bs].reshape(bs, bs)
                blk += 4 * np.eye(bs)
    A = to_scipy(n, bs, rowptr, colidx, nz)
    ilu = O.ILU0(n, bs, rowptr, colidx)
    assert ilu.factor(nz) == 0
    b = rng.standard_normal(n * bs)
    x = ilu.solve(b)
    assert np.allclose(A @ x, b, rtol=1e-11, atol=1e-11)


def test_ilu0_residual_zero_on_pattern(O, J):
    """Defining property of ILU(0): (L*U - A) vanishes on the sparsity pattern of A."""
    w, s, M0, p, nz, r = _twophase_system(O, J, dims=(5, 4, 3))
    n, bs = w["nc"], 2
    ilu = O.ILU0(n, bs, s["rowptr"], s["colidx"])
    assert ilu.factor(nz) == 0
    f = ilu.get()
    import scipy.sparse as sp
    L = to_scipy(n, bs, f["Lptr"], f["Lcol"], f["L"]) + sp.identity(n * bs)
    U = to_scipy(n, bs, f["Uptr"], f["Ucol"], f["U"])
    Dinv = f["D"].reshape(n, bs, bs).transpose(0, 2, 1)
    D = sp.block_diag([sp.coo_matrix(np.linalg.inv(d)) for d in Dinv])
    A = to_scipy(n, bs, s["rowptr"], s["colidx"], nz)
    E = (L @ (D + U) - A).toarray()
    mask = (to_scipy(n, bs, s["rowptr"], s["colidx"], np.ones_like(nz)).toarray() != 0)
    assert np.abs(E[mask]).max() <= 1e-12 * np.abs(A).max()


def test_block_jacobi_ilu_equals_per_block(O, J):
    """ilu0_csr(A, partition): each block is the ILU(0) of its own diagonal sub-matrix (par_ilu0.jl:47-87)."""
    w, s, M0, p, nz, r = _twophase_system(O, J, dims=(5, 4, 3))
    n, bs = w["nc"], 2
    part = O.partition_linear(3, n)
    ilu = O.ILU0(n, bs, s["rowptr"], s["colidx"], part)
    ilu.factor(nz)
    b = np.random.default_rng(5).standard_normal(n * bs)
    x = ilu.solve(b)
    A = to_scipy(n, bs, s["rowptr"], s["colidx"], nz).tocsr()
    for blk in range(1, 4):
        rows = np.where(part == blk)[0]
        dof = (rows[:, None] * bs + np.arange(bs)[None, :]).ravel()
        sub = A[dof][:, dof].tocoo()
        # sub-matrix as its own block CSR
        Ib = sub.row // bs + 1; Jb = sub.col // bs + 1
        rp, ci = O.csr_from_coo(Ib, Jb, rows.shape[0])
        dense = sub.toarray()
        nzb = np.concatenate([dense[(r_ * bs):(r_ * bs + bs), ((c_ - 1) * bs):((c_ - 1) * bs + bs)].T.ravel()
                              for r_ in range(rows.shape[0]) for c_ in ci[rp[r_] - 1:rp[r_ + 1] - 1]])
        sub_ilu = O.ILU0(rows.shape[0], bs, rp, ci)
        sub_ilu.factor(nzb)
        assert np.allclose(sub_ilu.solve(b[dof]), x[dof], rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("side", ["right", "left", "none"])
def test_bicgstab_solves(O, J, side):
    w, s, M0, p, nz, r = _twophase_system(O, J)
    n = w["nc"]
    O.scale_diagonal(n, 2, s["rowptr"], s["colidx"], nz, r)   # krylov scaling = :diagonal, keeps the unpreconditioned case solvable
    A = to_scipy(n, 2, s["rowptr"], s["colidx"], nz)
    ilu = O.ILU0(n, 2, s["rowptr"], s["colidx"]); ilu.factor(nz)
    x, st, its, hist = O.bicgstab(n, 2, s["rowptr"], s["colidx"], nz, r, ilu if side != "none" else None, side=side,
                                  rtol=1e-10, itmax=2000)
    assert st == 0 and its >= 1 and hist.shape[0] == its + 1
    xd = spla.spsolve(A.tocsc(), r)
    assert np.linalg.norm(x - xd) <= 1e-6 * np.linalg.norm(xd)
    if side != "left":   # history is the true residual norm for right/no preconditioning
        assert abs(np.linalg.norm(r - A @ x) - hist[-1]) <= 1e-6 * hist[0]


@pytest.mark.parametrize("side", ["right", "left"])
def test_bicgstab_history_equals_independent_recurrence(O, J, side):
    """The oracle's BiCGStab against an independent numpy transcription of the published recurrence (van der Vorst's
    preconditioned BiCGStab in the operation order SURVEY §8(c) records for Krylov.jl 0.9: x0 = 0, c = b, rho' = <c,r>, ...):
    the residual histories agree to rounding over the first iterations and the iteration counts are equal. Krylov.jl itself is
    not vendored, so this pins the restatement to the algorithm, not to a Julia run."""
    w, s, M0, p, nz, r = _twophase_system(O, J)
    n = w["nc"]
    O.scale_diagonal(n, 2, s["rowptr"], s["colidx"], nz, r)
    A = to_scipy(n, 2, s["rowptr"], s["colidx"], nz)
    ilu = O.ILU0(n, 2, s["rowptr"], s["colidx"]); ilu.factor(nz)
    rtol, atol = 1e-8, 1e-12
    x, st, its, hist = O.bicgstab(n, 2, s["rowptr"], s["colidx"], nz, r, ilu, side=side, rtol=rtol, atol=atol, itmax=500)
    P = ilu.solve
    N = P if side == "right" else (lambda v: v)
    M = P if side == "left" else (lambda v: v)
    b = r
    rr = M(b).copy(); c = b.copy()                       # Krylov.jl: c = b also with a left preconditioner
    pp = rr.copy(); xx = np.zeros_like(b)
    rho_next = c @ rr
    h = [np.linalg.norm(rr)]
    eps = atol + rtol * h[0]
    for _ in range(500):
        rho = rho_next
        y = N(pp); v = M(A @ y)
        alpha = rho / (c @ v)
        sv = rr - alpha * v
        xx = xx + alpha * y
        z = N(sv); t = M(A @ z)
        omega = (t @ sv) / (t @ t)
        xx = xx + omega * z
        rr = sv - omega * t
        rho_next = c @ rr
        beta = (rho_next / rho) * (alpha / omega)
        pp = rr + beta * (pp - omega * v)
        h.append(np.linalg.norm(rr))
        if h[-1] <= eps:
            break
    m = min(8, len(h), len(hist))
    assert np.allclose(hist[:m], h[:m], rtol=1e-8)
    assert abs(its - (len(h) - 1)) <= max(1, its // 20)
    # Both iterates satisfy the stopping rule, so they agree in the norm the rule uses (the preconditioned residual). The plain
    # difference is only bounded by cond(A) * rtol and moves with the summation order of the OpenMP inner products, hence the
    # looser second bound.
    assert np.linalg.norm(M(A @ (x - xx))) <= 4 * eps
    assert np.linalg.norm(x - xx) <= 1e-4 * np.linalg.norm(xx)


def test_bicgstab_min_iterations_and_itmax(O, J):
    w, s, M0, p, nz, r = _twophase_system(O, J)
    n = w["nc"]
    ilu = O.ILU0(n, 2, s["rowptr"], s["colidx"]); ilu.factor(nz)
    _, st, its, _ = O.bicgstab(n, 2, s["rowptr"], s["colidx"], nz, r, ilu, rtol=1e-1, min_it=1)
    _, st2, its2, _ = O.bicgstab(n, 2, s["rowptr"], s["colidx"], nz, r, ilu, rtol=1e-1, min_it=its + 3)
    assert st == 0 and st2 == 0 and its2 >= its + 3 - 1
    _, st3, its3, _ = O.bicgstab(n, 2, s["rowptr"], s["colidx"], nz, r, None, side="none", rtol=1e-14, itmax=2)
    assert st3 == 1 and its3 == 2
    x, st4, its4, h4 = O.bicgstab(n, 2, s["rowptr"], s["colidx"], nz, np.zeros_like(r), ilu)
    assert st4 == 0 and its4 == 0 and np.all(x == 0)


def test_update_rules(O):
    # choose_increment order: scale, abs, rel, lower, upper (src/variables/utils.jl:146-156)
    v = np.array([1.0, 1.0, 1.0, 0.05, 0.95])
    dx = np.array([10.0, -10.0, 0.3, -1.0, 1.0])
    out = O.update_scalar(v.copy(), dx, abs_max=0.5, rel_max=0.2, minv=0.0, maxv=1.0)
    assert np.allclose(out, [1.0, 0.8, 1.0, 0.04, 1.0])
    # unit_update_pairs!: w applied after limiting; s2 mirrors s1
    s = np.array([0.5, 0.5, 0.1, 0.9, 0.9, 0.1])
    O.update_fraction_pair(s, np.array([0.4, -0.5, 0.5]), abs_max=0.2)
    assert np.allclose(s, [0.7, 0.3, 0.0, 1.0, 1.0, 0.0])
    assert O.increment_norm(np.array([1.0, -3.0, 2.0])) == (6.0, 3.0)
    assert O.maxabs_rows(np.array([1.0, -2.0, -5.0, 0.5]), 2).tolist() == [5.0, 2.0]


def test_scaling(O, J):
    w, s, M0, p, nz, r = _twophase_system(O, J, dims=(4, 3, 2))
    n = w["nc"]
    A = to_scipy(n, 2, s["rowptr"], s["colidx"], nz)
    x = spla.spsolve(A.tocsc(), r)
    nz2, r2 = nz.copy(), r.copy()
    O.scale_diagonal(n, 2, s["rowptr"], s["colidx"], nz2, r2)
    A2 = to_scipy(n, 2, s["rowptr"], s["colidx"], nz2)
    assert np.allclose(spla.spsolve(A2.tocsc(), r2), x, rtol=1e-8)
    d = A2.toarray()
    for c in range(n):
        assert np.allclose(d[2 * c:2 * c + 2, 2 * c:2 * c + 2], np.eye(2), atol=1e-8)


@pytest.mark.parametrize("side,restart", [("right", False), ("left", False), ("none", False), ("right", True)])
def test_gmres_solves(O, J, side, restart):
    """Krylov.jl gmres! restatement: monotone residual estimates, true residual matches the estimate (right/none)."""
    w, s, M0, p, nz, r = _twophase_system(O, J)
    n = w["nc"]
    O.scale_diagonal(n, 2, s["rowptr"], s["colidx"], nz, r)
    A = to_scipy(n, 2, s["rowptr"], s["colidx"], nz)
    ilu = O.ILU0(n, 2, s["rowptr"], s["colidx"]); ilu.factor(nz)
    x, st, its, hist = O.gmres(n, 2, s["rowptr"], s["colidx"], nz, r, ilu if side != "none" else None, side=side, rtol=1e-9,
                               itmax=400, memory=20, restart=restart)
    assert st == 0 and its >= 1
    if not restart:
        assert hist.shape[0] == its + 1 and np.all(np.diff(hist) <= 1e-14 * hist[0])     # GMRES residuals never increase
    xd = spla.spsolve(A.tocsc(), r)
    assert np.linalg.norm(x - xd) <= 1e-5 * np.linalg.norm(xd)
    if side != "left":
        assert abs(np.linalg.norm(r - A @ x) - hist[-1]) <= 1e-6 * hist[0]


def test_gmres_history_is_the_minimal_residual_over_the_krylov_space(O, J):
    """Independent pin of the oracle's GMRES (Krylov.jl order): for right preconditioning the j-th entry of stats.residuals is
    min over x in N^-1 K_j(A N^-1, b) of |b - A x| — computed here by dense least squares on an explicitly orthonormalised
    Krylov basis, without Arnoldi recurrences or Givens rotations."""
    w, s, M0, p, nz, r = _twophase_system(O, J, dims=(5, 4, 3))
    n = w["nc"]
    O.scale_diagonal(n, 2, s["rowptr"], s["colidx"], nz, r)
    A = to_scipy(n, 2, s["rowptr"], s["colidx"], nz)
    ilu = O.ILU0(n, 2, s["rowptr"], s["colidx"]); ilu.factor(nz)
    x, st, its, hist = O.gmres(n, 2, s["rowptr"], s["colidx"], nz, r, ilu, side="right", rtol=1e-12, itmax=12, memory=20, restart=False)
    op = lambda v: A @ ilu.solve(v)
    K = [r / np.linalg.norm(r)]
    for j in range(1, min(8, len(hist))):
        # orthonormal basis of K_j by repeated full orthogonalisation (numerically independent of the oracle's modified Gram-Schmidt)
        W_ = np.column_stack([op(q) for q in K])
        y, *_ = np.linalg.lstsq(W_, r, rcond=None)
        assert np.isclose(np.linalg.norm(r - W_ @ y), hist[j], rtol=1e-6), j
        v = op(K[-1])
        Q = np.column_stack(K)
        for _ in range(2):
            v = v - Q @ (Q.T @ v)
        K.append(v / np.linalg.norm(v))


@pytest.mark.parametrize("restart", [False, True])
def test_fgmres_equals_right_preconditioned_gmres(O, J, restart):
    """Krylov.jl fgmres! with a FIXED preconditioner spans the same spaces as right-preconditioned gmres!: same residual
    history to rounding, same solution; the restarted distributed form (left preconditioning, restart = true,
    ext/JutulPartitionedArraysExt/krylov.jl:67-74) solves the same system in more iterations."""
    w, s, M0, p, nz, r = _twophase_system(O, J)
    n = w["nc"]
    O.scale_diagonal(n, 2, s["rowptr"], s["colidx"], nz, r)
    A = to_scipy(n, 2, s["rowptr"], s["colidx"], nz)
    ilu = O.ILU0(n, 2, s["rowptr"], s["colidx"]); ilu.factor(nz)
    kw = dict(rtol=1e-7, itmax=600, memory=40, restart=restart)
    xg, stg, itg, hg = O.gmres(n, 2, s["rowptr"], s["colidx"], nz, r, ilu, side="right", **kw)
    xf, stf, itf, hf = O.gmres(n, 2, s["rowptr"], s["colidx"], nz, r, ilu, side="right", flexible=True, **kw)
    assert stg == 0 and stf == 0 and abs(itg - itf) <= 1
    m = min(len(hg), len(hf))
    assert np.allclose(hg[:m], hf[:m], rtol=1e-6, atol=1e-12 * hg[0])
    assert np.linalg.norm(xg - xf) <= 1e-7 * np.linalg.norm(xg)
    xl, stl, itl, hl = O.gmres(n, 2, s["rowptr"], s["colidx"], nz, r, ilu, side="left", rtol=1e-7, itmax=600, memory=40, restart=True)
    xd = spla.spsolve(A.tocsc(), r)
    assert stl == 0 and np.linalg.norm(xl - xd) <= 5e-5 * np.linalg.norm(xd)      # (error amplification ~1e2) x rtol
    if restart:
        x20, st20, it20, h20 = O.gmres(n, 2, s["rowptr"], s["colidx"], nz, r, ilu, side="right", rtol=1e-7, itmax=600, memory=600, restart=True)
        assert itg >= it20          # a short memory never beats the full basis


def _nfvm_case(rng, nc=50, nf=120, nph=2, max_mpfa=5):
    left = rng.integers(1, nc + 1, nf); right = (left + rng.integers(0, nc - 1, nf)) % nc + 1

    def half():
        cnt = rng.integers(0, max_mpfa + 1, nf)
        ptr = np.concatenate([[1], 1 + np.cumsum(cnt)])
        return dict(T_left=rng.standard_normal(nf), T_right=rng.standard_normal(nf), ptr=ptr, cell=rng.integers(1, nc + 1, ptr[-1] - 1),
                    T=rng.standard_normal(ptr[-1] - 1))
    return left, right, half(), half(), rng.standard_normal(nph * nc)


def test_nfvm_evaluate_flux_oracle(O):
    """src/NFVM/evaluation.jl: linear scheme with an empty MPFA remainder is the TPFA flux T_l p_l + T_r p_r; the
    nonlinear blend reduces to 0.5 (q_l - q_r) when both remainders vanish."""
    rng = np.random.default_rng(0)
    left, right, L, R, p = _nfvm_case(rng, nph=1)
    q = O.nfvm_evaluate_flux(left, right, L, None, p, scheme="linear")
    ref = np.array([L["T_left"][f] * p[left[f] - 1] + L["T_right"][f] * p[right[f] - 1]
                    + sum(p[L["cell"][k - 1] - 1] * L["T"][k - 1] for k in range(L["ptr"][f], L["ptr"][f + 1])) for f in range(left.shape[0])])
    assert np.allclose(q, ref, rtol=1e-14, atol=1e-14)
    empty = lambda d: dict(d, ptr=np.ones_like(d["ptr"]), cell=np.zeros(0, dtype=np.int64), T=np.zeros(0))
    Le, Re = empty(L), empty(R)
    qn = O.nfvm_evaluate_flux(left, right, Le, Re, p, scheme="ntpfa")
    ql = Le["T_left"] * p[left - 1] + Le["T_right"] * p[right - 1]
    qr = -(Re["T_left"] * p[left - 1] + Re["T_right"] * p[right - 1])
    assert np.allclose(qn, 0.5 * ql - 0.5 * qr, rtol=1e-14, atol=1e-14)
    # the weights sum to one: mu_l + mu_r = 1 => q lies between the two half fluxes for :nmpfa
    qm = O.nfvm_evaluate_flux(left, right, L, R, p, scheme="nmpfa")
    rl = np.array([sum(p[L["cell"][k - 1] - 1] * L["T"][k - 1] for k in range(L["ptr"][f], L["ptr"][f + 1])) for f in range(left.shape[0])])
    rr = -np.array([sum(p[R["cell"][k - 1] - 1] * R["T"][k - 1] for k in range(R["ptr"][f], R["ptr"][f + 1])) for f in range(left.shape[0])])
    q_l = L["T_left"] * p[left - 1] + L["T_right"] * p[right - 1] + rl
    q_r = -(R["T_left"] * p[left - 1] + R["T_right"] * p[right - 1]) + rr
    tot = np.abs(rl) + np.abs(rr)
    mu_l = np.where(tot < 1e-10, 0.5, np.abs(rr) / np.where(tot == 0, 1, tot)); mu_r = np.where(tot < 1e-10, 0.5, np.abs(rl) / np.where(tot == 0, 1, tot))
    assert np.allclose(qm, mu_l * q_l - mu_r * q_r, rtol=1e-13, atol=1e-13)


def test_diagonal_preconditioners(O, J):
    """Jacobi (w * inv(A_ii)) and SPAI(0) (inv(|row|_F^2) * A_ii) as the reference defines them
    (src/linsolve/precond/jacobi.jl:15-18, spai.jl:42-62); both are exact inverses of a (block-)diagonal matrix up to
    their definition: Jacobi with w = 1 inverts block-diagonal A, SPAI(0) of a scalar diagonal matrix is 1/a."""
    w, s, M0, p, nz, r = _twophase_system(O, J, dims=(4, 3, 2))
    n = w["nc"]
    A = to_scipy(n, 2, s["rowptr"], s["colidx"], nz).toarray()
    D = O.jacobi_factor(n, 2, s["rowptr"], s["colidx"], nz, w=1.0)
    y = np.random.default_rng(2).standard_normal(2 * n)
    x = O.diagonal_apply(D, y, 2)
    for c in range(n):
        blk = A[2 * c:2 * c + 2, 2 * c:2 * c + 2]
        assert np.allclose(blk @ x[2 * c:2 * c + 2], y[2 * c:2 * c + 2], rtol=1e-9)
    D23 = O.jacobi_factor(n, 2, s["rowptr"], s["colidx"], nz)
    assert np.allclose(D23, (2.0 / 3.0) * D)
    # SPAI(0): row-wise minimiser of |I - M A|_F over diagonal-block-scaled M = d_i * A_ii / ... (definition check)
    S = O.spai0_factor(n, 2, s["rowptr"], s["colidx"], nz)
    for c in range(n):
        row = A[2 * c:2 * c + 2, :]
        assert np.allclose(S.reshape(n, 4)[c].reshape(2, 2).T, row[:, 2 * c:2 * c + 2] / (row ** 2).sum(), rtol=1e-12)
    # scalar diagonal matrix: SPAI(0) = 1/a
    rp = np.arange(1, 6, dtype=np.int64); ci = np.arange(1, 5, dtype=np.int64); a = np.array([2.0, -4.0, 0.5, 10.0])
    assert np.allclose(O.spai0_factor(4, 1, rp, ci, a), 1.0 / a)
    assert np.allclose(O.diagonal_apply(O.jacobi_factor(4, 1, rp, ci, a, w=1.0), a, 1), np.ones(4))


def test_two_colour_operator_identity(O, J):
    """The claim behind the identity rows of the device SpMV (csrc/krylov.cu): with a two-colour numbering and ILU(0),
    x = N^{-1} w satisfies (A x)_i = w_i on every first-colour row (no L entries, all couplings kept in U), because
    U_ij = A_ij and D_i = A_ii there. Checked with the oracle's own ILU(0) and SpMV on the renumbered system; the
    deviation is rounding, amplified by the conditioning of the 2x2 diagonal blocks."""
    w = J.workloads.unstructured_hex(9, 8, 7)
    n = w["nc"]
    perm, ncol = J.multicolor_ordering(w["N"], n)
    assert ncol == 2
    w2 = dict(w)
    w2["N"] = perm[w["N"] - 1]
    for k in ("pv", "p0", "sw0"):
        a = np.empty_like(w[k]); a[perm - 1] = w[k]; w2[k] = a
    w2["src_cells"] = perm[w["src_cells"] - 1]
    s = oracle_system(O, w2)
    M0 = O.mass_2ph(w2["pv"], w2["params"], w2["p0"], w2["sw0"])
    rng = np.random.default_rng(8)
    p = w2["p0"] * (1 + 1e-3 * rng.standard_normal(n))
    nz, r = O.assemble_2ph(s["hf"], s["diag_pos"], s["hf_pos"], w2["Tf"], w2["gdz"], w2["pv"], w2["params"], p, w2["sw0"], M0, w2["dt"],
                           s["colidx"].shape[0], w2["src_cells"], w2["src_vals"])
    ilu = O.ILU0(n, 2, s["rowptr"], s["colidx"]); ilu.factor(nz)
    f = ilu.get()
    first = np.diff(f["Lptr"]) == 0                      # rows without L entries = first colour
    assert first.sum() >= n // 2 - 1 and np.all(np.diff(f["Uptr"])[first] == np.diff(s["rowptr"])[first] - 1)
    wv = rng.standard_normal(2 * n)
    x = ilu.solve(wv)
    Ax = O.spmv(n, 2, s["rowptr"], s["colidx"], nz, x)
    rows = np.repeat(first, 2)
    A = to_scipy(n, 2, s["rowptr"], s["colidx"], nz)
    scale = (abs(A) @ np.abs(x))[rows]                   # |A||x|: what rounding in the product is relative to
    assert np.all(np.abs(Ax[rows] - wv[rows]) <= 1e-12 * scale + 1e-300)
    assert np.abs(Ax[~rows] - wv[~rows]).max() > 1e-3 * np.abs(wv).max()      # the second colour is NOT an identity


def test_schur_complement_reduction(O, J):
    """MultiModel :schur_apply (src/linsolve/multimodel.jl:17-160): reduce the right-hand side, solve with the operator
    B - C E^-1 D, recover the eliminated unknowns — equals the direct solve of the full 2x2 block system, with Jutul's
    sign convention dx = -(J \\ r)."""
    w, s, M0, p, nz, r = _twophase_system(O, J, dims=(5, 4, 3))
    n = w["nc"]
    B = to_scipy(n, 2, s["rowptr"], s["colidx"], nz).tocsr()
    rng = np.random.default_rng(12)
    sizes = [3, 5]                                          # two eliminated groups (e.g. two wells)
    sc = abs(B).max()
    C = [sc * 1e-2 * rng.standard_normal((2 * n, m)) * (rng.random((2 * n, m)) < 0.05) for m in sizes]
    D = [sc * 1e-2 * rng.standard_normal((m, 2 * n)) * (rng.random((m, 2 * n)) < 0.05) for m in sizes]
    E = [sc * (np.eye(m) + 0.1 * rng.standard_normal((m, m))) for m in sizes]
    a = r.copy(); b = [sc * rng.standard_normal(m) for m in sizes]
    import scipy.sparse as sp
    full = sp.bmat([[B, sp.csr_matrix(C[0]), sp.csr_matrix(C[1])], [sp.csr_matrix(D[0]), sp.csr_matrix(E[0]), None],
                    [sp.csr_matrix(D[1]), None, sp.csr_matrix(E[1])]]).tocsc()
    z = spla.spsolve(full, np.concatenate([a] + b))
    a_red = O.schur_prepare(a.copy(), C, E, b)
    S = spla.LinearOperator((2 * n, 2 * n), matvec=lambda v: O.schur_mul(B, C, D, E, v))
    dx_solver, info = spla.gmres(S, a_red, rtol=1e-13, atol=0.0, restart=200, maxiter=5)
    assert info == 0
    x, y = O.schur_dx_update(D, E, b, dx_solver)
    assert np.linalg.norm(x + z[: 2 * n]) <= 1e-8 * np.linalg.norm(z[: 2 * n])
    off = 2 * n
    for yi, m in zip(y, sizes):
        assert np.linalg.norm(yi + z[off: off + m]) <= 1e-8 * max(np.linalg.norm(z[off: off + m]), 1e-30)
        off += m
    # alpha / beta form of the operator
    v = rng.standard_normal(2 * n); res0 = rng.standard_normal(2 * n)
    assert np.allclose(O.schur_mul(B, C, D, E, v, alpha=2.0, beta=-0.5, res=res0), 2.0 * O.schur_mul(B, C, D, E, v) - 0.5 * res0, rtol=1e-12)


def test_level_scheduled_solve_is_bitwise_the_sequential_sweep(J, O):
    """The checker's optional level schedule of ilu_solve! (rows of a dependency level concurrently) performs, row by row, the
    operations of the sequential sweep (src/StaticCSR/ilu0.jl:156-187): x must be bitwise identical, for a multicolour
    numbering (2 levels) and for the mesh's own numbering (many levels)."""
    w = J.workloads.unstructured_hex(14, 12, 10)
    n = w["nc"]
    for relabel in (True, False):
        wp = dict(w)
        if relabel:
            perm, ncol = J.multicolor_ordering(w["N"], n)
            wp["N"] = perm[w["N"] - 1]
        s = oracle_system(O, wp)
        rng = np.random.default_rng(5)
        nz = rng.standard_normal(s["colidx"].shape[0] * 4).reshape(-1, 4)
        nz[s["diag_pos"] - 1] += np.array([25.0, 0.0, 0.0, 25.0])
        ilu = O.ILU0(n, 2, s["rowptr"], s["colidx"])
        assert ilu.factor(nz.ravel()) == 0
        b = rng.standard_normal(2 * n)
        x_seq = ilu.solve(b)
        nlev = ilu.set_level_schedule(True, max_levels=10000)
        assert nlev == (2 if relabel else nlev) and nlev >= 2
        assert np.array_equal(ilu.solve(b), x_seq)
        assert ilu.set_level_schedule(True, max_levels=1) == 0          # refused: too many levels, sequential sweep stays
        assert np.array_equal(ilu.solve(b), x_seq)
