"""CPU pins of the oracle restatements behind the widened rows (oracle/widen.py): interpolation against the reference's
own known answers (test/utils.jl:154-246), the variable ordering, the NFVM flux partials (finite differences and the C++
value oracle), the :fvm assembly and the adjoint alignment (transposition identity)."""
import numpy as np
import pytest

from oracle import widen as W


@pytest.mark.parametrize("constant_dx", [True, False, None])
def test_get_1d_interpolator_reference_known_answers(constant_dx):
    """test/utils.jl:154-178"""
    x = np.arange(0, 4.0001, 0.1)
    I = W.get_1d_interpolator(x, np.sin(x), constant_dx=constant_dx)
    assert abs(I(np.pi / 2) - 1.0) <= 1e-2
    x = [0.0, 0.5, 1.0]
    F = W.get_1d_interpolator(x, np.square(x), constant_dx=constant_dx)
    assert np.isclose(F(0.5), 0.25) and np.isclose(F(0.25), 0.25 / 2)
    # default is to not extrapolate (block values of the reference test, component by component)
    for comp, (x1, x2, x3) in enumerate([(0.0, 0.25, 1.0), (0.1, 0.35, 1.1)]):
        Fb = W.get_1d_interpolator(x, [x1, x2, x3], constant_dx=constant_dx)
        assert np.isclose(Fb(0.5), x2) and np.isclose(Fb(0.25), (x1 + x2) / 2)
        assert np.isclose(Fb(-1.0), x1) and np.isclose(Fb(1.5), x3)


@pytest.mark.parametrize("cdx", [True, False, None])
@pytest.mark.parametrize("cdy", [True, False, None])
def test_get_2d_interpolator_reference_known_answers(cdx, cdy):
    """test/utils.jl:180-230: nodes are reproduced, ForwardDiff gradient == finite differences on a fine grid."""
    f = lambda x, y: np.sin(x) + np.cos(y) + 0.5 * x
    xs = np.linspace(0.0, 4.0, 10); ys = np.linspace(0.0, 5.0, 8)
    fs = f(xs[:, None], ys[None, :])
    I = W.get_2d_interpolator(xs, ys, fs, constant_dx=cdx, constant_dy=cdy)
    for x in xs:
        for y in ys:
            assert np.isclose(I(x, y), f(x, y))
    eps = 1e-6
    for x in np.linspace(-1.0, 5.0, 100)[::4]:        # the reference's fine grid (no sample sits on a kink), every 4th point
        for y in np.linspace(-1.0, 6.0, 100)[::4]:
            r = I(W.Dual(x, np.array([1.0, 0.0])), W.Dual(y, np.array([0.0, 1.0])))
            v0 = I(x, y)
            assert np.isclose(r.v, v0)
            assert np.isclose((I(x + eps, y) - v0) / eps, r.d[0], rtol=1e-3, atol=1e-8)
            assert np.isclose((I(x, y + eps) - v0) / eps, r.d[1], rtol=1e-3, atol=1e-8)


def test_first_lower_fast_lookup_reference_property():
    """test/utils.jl:232-249"""
    for nstep in range(2, 101, 7):
        for start, stop in [(1.5, 3.9), (-100.0, 53.0), (-1e-3, 1e-3)]:
            dx = np.linspace(start, stop, nstep)
            lookup = W.interpolation_constant_lookup(dx)
            assert lookup is not None and np.isclose(lookup[1], (stop - start) / (nstep - 1))
            wd = 0.1 * (stop - start)
            for x in np.linspace(start - wd, stop + wd, 3 * nstep):
                pos = W.first_lower(dx, x)
                pos_l = W.first_lower(dx, x, lookup)
                at_boundary = abs(x - (start + pos * lookup[1])) <= 1e-10 or abs(x - (start + pos_l * lookup[1])) <= 1e-10
                assert pos == pos_l or at_boundary


def test_sort_secondary_variables_properties():
    prim = ["Pressure", "Saturation"]
    par = ["PoreVolume"]
    sec = {"Mass": ["Density", "So", "PoreVolume"], "Mobility": ["Kr", "Viscosity"], "Kr": ["So"], "So": ["Saturation"],
           "Density": ["Pressure"], "Viscosity": ["Pressure"]}
    order = W.sort_secondary_variables(prim, sec, par)
    assert sorted(order) == sorted(sec)
    for k, deps in sec.items():
        for d in deps:
            if d in sec:
                assert order.index(d) < order.index(k)
    # depth-first post-order in declaration order, dependencies in ascending node order: Mass pulls So (node 7) and Density
    # (node 8) first, then Mobility pulls Kr and Viscosity
    assert order == ["So", "Density", "Mass", "Kr", "Viscosity", "Mobility"]
    with pytest.raises(ValueError):
        W.sort_secondary_variables(prim, {"A": ["B"], "B": ["A"]}, par)
    with pytest.raises(KeyError):
        W.sort_secondary_variables(prim, {"A": ["Missing"]}, par)
    with pytest.raises(ValueError):
        W.sort_secondary_variables(prim, {"Pressure": []}, par)


def _random_nfvm(rng, nc, nf, nm=3):
    left = rng.integers(1, nc + 1, nf); right = rng.integers(1, nc + 1, nf)
    right = np.where(right == left, right % nc + 1, right)

    def half():
        cnt = rng.integers(0, nm + 1, nf)
        ptr = np.concatenate([[1], 1 + np.cumsum(cnt)]).astype(np.int64)
        return dict(T_left=rng.uniform(0.5, 2.0, nf), T_right=-rng.uniform(0.5, 2.0, nf), ptr=ptr,
                    cell=rng.integers(1, nc + 1, ptr[-1] - 1).astype(np.int64), T=rng.normal(0, 0.3, ptr[-1] - 1))
    return left.astype(np.int64), right.astype(np.int64), half(), half()


@pytest.mark.parametrize("scheme", ["linear", "ntpfa", "nmpfa"])
def test_nfvm_dual_flux_matches_value_oracle_and_fd(O, scheme):
    rng = np.random.default_rng(5)
    nc, nf = 40, 90
    left, right, L, R = _random_nfvm(rng, nc, nf)
    p = rng.uniform(1.0, 2.0, nc)
    q_ref = O.nfvm_evaluate_flux(left, right, L, None if scheme == "linear" else R, p, scheme=scheme)
    vpos, vars_ = W.nfvm_discretization_stencil(left, right, L, None if scheme == "linear" else R)
    for f in range(nf):
        qf = W.nfvm_evaluate_flux_dual(f, left, right, L, R, scheme, lambda c: p[c - 1])
        assert np.isclose(qf, q_ref[f], rtol=1e-14, atol=1e-14)
        for s in range(vpos[f] - 1, vpos[f + 1] - 1):
            var = int(vars_[s])
            d = W.nfvm_evaluate_flux_dual(f, left, right, L, R, scheme, lambda c: W.Dual(p[c - 1], np.array([1.0 if c == var else 0.0])))
            h = 1e-6
            pp = p.copy(); pp[var - 1] += h
            pm = p.copy(); pm[var - 1] -= h
            fd = (W.nfvm_evaluate_flux_dual(f, left, right, L, R, scheme, lambda c: pp[c - 1]) -
                  W.nfvm_evaluate_flux_dual(f, left, right, L, R, scheme, lambda c: pm[c - 1])) / (2 * h)
            assert np.isclose(d.d[0], fd, rtol=1e-5, atol=1e-7)


def test_fvm_assembly_jacobian_is_derivative_of_residual(O):
    rng = np.random.default_rng(11)
    nc, nf = 25, 60
    left, right, L, R = _random_nfvm(rng, nc, nf)
    vpos, vars_ = W.nfvm_discretization_stencil(left, right, L, R)
    I, Jc = W.fvm_declare_pattern(nc, left, right, vpos, vars_)
    rowptr, colidx = O.csr_from_coo(I, Jc, nc)
    lp, rp = W.fvm_align(left, right, vpos, vars_, rowptr, colidx)
    dpos = np.array([W.find_jac_position_block(rowptr, colidx, c, c, 1, 1, 1) for c in range(1, nc + 1)])
    p = rng.uniform(1.0, 2.0, nc)
    acc = rng.normal(size=nc); dacc = rng.uniform(1, 2, nc)
    nz, r, q = W.fvm_assemble_nfvm(nc, left, right, L, R, "ntpfa", p, vpos, vars_, lp, rp, dpos, len(colidx), acc, dacc)
    import scipy.sparse as sp
    Jm = sp.csr_matrix((nz, colidx - 1, rowptr - 1), shape=(nc, nc)).toarray()
    h = 1e-6
    for c in range(nc):
        pp = p.copy(); pp[c] += h
        pm = p.copy(); pm[c] -= h
        rp_ = W.fvm_assemble_nfvm(nc, left, right, L, R, "ntpfa", pp, vpos, vars_, lp, rp, dpos, len(colidx))[1]
        rm_ = W.fvm_assemble_nfvm(nc, left, right, L, R, "ntpfa", pm, vpos, vars_, lp, rp, dpos, len(colidx))[1]
        col = (rp_ - rm_) / (2 * h)
        col[c] += dacc[c]
        assert np.allclose(Jm[:, c], col, rtol=1e-5, atol=1e-6)
    # mass balance of the flux part: every face adds q to one cell and removes it from the other
    assert abs((r - acc).sum()) <= 1e-12 * np.abs(q).sum()


def test_adjoint_alignment_is_block_transposition(O):
    """as_adjoint positions (src/equations.jl:101-108): writing entry (row, col, eq, partial) at the adjoint position of the
    transposed pattern produces exactly J^T."""
    rng = np.random.default_rng(3)
    n, N = 7, 2
    I = np.concatenate([np.arange(1, n + 1), rng.integers(1, n + 1, 15)]); Jc = np.concatenate([np.arange(1, n + 1), rng.integers(1, n + 1, 15)])
    rowptr, colidx = O.csr_from_coo(I, Jc, n)
    rowptrT, colidxT = O.csr_from_coo(Jc, I, n)
    nzb = len(colidx)
    nz = np.zeros(nzb * N * N); nzT = np.zeros(nzb * N * N)
    for row in range(1, n + 1):
        for k in range(rowptr[row - 1] - 1, rowptr[row] - 1):
            col = colidx[k]
            for eq in (1, 2):
                for d in (1, 2):
                    v = rng.normal()
                    nz[W.find_jac_position_block(rowptr, colidx, row, col, eq, d, N) - 1] = v
                    nzT[W.find_jac_position_block(rowptrT, colidxT, row, col, eq, d, N, adjoint=True) - 1] = v
    from conftest import to_scipy
    A = to_scipy(n, N, rowptr, colidx, nz).toarray()
    AT = to_scipy(n, N, rowptrT, colidxT, nzT).toarray()
    assert np.array_equal(AT, A.T)


def test_secondary_variable_graph_against_closed_forms():
    """update_secondary_variables_state! on Duals reproduces hand-written derivatives of a density / mobility / mass chain."""
    rng = np.random.default_rng(2)
    nc = 50
    p = rng.uniform(1e5, 3e5, nc); sw = rng.uniform(0.05, 0.95, nc); pv = rng.uniform(1, 2, nc)
    mu_tab = W.get_1d_interpolator(np.linspace(5e4, 4e5, 8), 1e-3 * (1 + 1e-6 * np.linspace(5e4, 4e5, 8)))
    defs = {
        "Density": dict(kind="exp", deps=["Pressure"], c=[1000.0, 4.5e-10, 1e5]),
        "Kr": dict(kind="power", deps=["Saturation"], c=[0.9, 2.0, 0.1, 0.8]),
        "Viscosity": dict(kind="table1d", deps=["Pressure"], c=[1.0], table=mu_tab),
        "Mobility": dict(kind="quotient", deps=["Kr", "Viscosity"], c=[1.0]),
        "Mass": dict(kind="product", deps=["Density", "Saturation", "PoreVolume"], c=[1.0]),
    }
    order = W.sort_secondary_variables(["Pressure", "Saturation"], {k: v["deps"] for k, v in defs.items()}, ["PoreVolume"])
    e = np.eye(2)
    state = {"Pressure": W.Dual(p, e[:, :1] * np.ones(nc)), "Saturation": W.Dual(sw, e[:, 1:] * np.ones(nc)), "PoreVolume": pv}
    W.update_secondary_variables_state(state, defs, order)
    rho = 1000.0 * np.exp(4.5e-10 * (p - 1e5))
    assert np.allclose(state["Density"].v, rho) and np.allclose(state["Density"].d[0], 4.5e-10 * rho) and np.all(state["Density"].d[1] == 0)
    u = np.clip((sw - 0.1) / 0.8, 0, 1)
    assert np.allclose(state["Kr"].v, 0.9 * u ** 2)
    inside = (sw >= 0.1) & (sw <= 0.9)
    assert np.allclose(state["Kr"].d[1], np.where(inside, 0.9 * 2 * u / 0.8, 0.0))
    assert np.allclose(state["Mass"].v, rho * sw * pv) and np.allclose(state["Mass"].d[1], rho * pv) and np.allclose(state["Mass"].d[0], 4.5e-10 * rho * sw * pv)
    mob = state["Mobility"]
    assert np.allclose(mob.v, state["Kr"].v / state["Viscosity"].v)
    h = 1.0
    mu_p = np.array([mu_tab(x + h) for x in p]); mu_m = np.array([mu_tab(x - h) for x in p])
    assert np.allclose(mob.d[0], -(state["Kr"].v / state["Viscosity"].v ** 2) * (mu_p - mu_m) / (2 * h), rtol=1e-6, atol=1e-12)


def test_property_dual_assembly_equals_closed_form_oracle(O, J):
    """fill_conservation_eq! on property Duals (oracle/widen.py) with the closed forms of the built-in law as properties
    reproduces the C++ restatement of the same law to rounding: the two derivations of the Jacobian agree."""
    from conftest import oracle_system
    w = J.workloads.unstructured_hex(6, 5, 4)
    nc = w["nc"]; sy = oracle_system(O, w)
    p = w["p0"] * (1 + 1e-3 * np.sin(np.arange(nc))); sw = w["sw0"]; pv = w["pv"]; par = w["params"]
    M0 = O.mass_2ph(pv, par, w["p0"], sw)
    rho = [par[0] * np.exp(par[2] * (p - par[6])), par[1] * np.exp(par[3] * (p - par[6]))]
    S = [sw, 1 - sw]; dS = [1.0, -1.0]
    props = {}
    for a, n in enumerate("WO"):
        mob = rho[a] * S[a] ** 2 / par[4 + a]
        props["Density" + n] = np.stack([rho[a], par[2 + a] * rho[a], 0 * p])
        props["Mobility" + n] = np.stack([mob, par[2 + a] * mob, rho[a] * 2 * S[a] * dS[a] / par[4 + a]])
        props["Mass" + n] = np.stack([pv * rho[a] * S[a], pv * par[2 + a] * rho[a] * S[a], pv * rho[a] * dS[a]])
    nnzb = sy["colidx"].shape[0]
    nz, r = W.assemble_2ph_props(sy["hf"], sy["diag_pos"], sy["hf_pos"], w["Tf"], w["gdz"], p, props, M0, w["dt"], nnzb, w["src_cells"], w["src_vals"])
    nz2, r2 = O.assemble_2ph(sy["hf"], sy["diag_pos"], sy["hf_pos"], w["Tf"], w["gdz"], pv, par, p, sw, M0, w["dt"], nnzb, w["src_cells"], w["src_vals"])
    assert np.abs(r - r2).max() <= 1e-14 * np.abs(r2).max() and np.abs(nz - nz2).max() <= 1e-14 * np.abs(nz2).max()


def test_two_restatements_of_the_update_rules_agree(O):
    """choose_increment / unit_update_pairs! exist twice in the checker (C++ in oracle.cpp, Python in oracle/widen.py, written
    at different times from src/variables/utils.jl:149-174,471-482): bit-equal on random inputs, and the nf = 2 case of the
    many-fraction update reduces to the pair update."""
    rng = np.random.default_rng(6)
    n = 500
    v = rng.uniform(0, 2, n); dx = rng.normal(0, 1, n)
    a = v.copy()
    O.update_scalar(a, dx, w=0.7, abs_max=0.3, rel_max=0.2, minv=0.1, maxv=1.9, scale=1.5)
    b = np.array([vi + W._choose_increment(vi, 0.7 * di, 0.3, 0.2, 0.1, 1.9, 1.5) for vi, di in zip(v, dx)])
    assert np.array_equal(a, b)
    sw = rng.uniform(0, 1, n)
    s = np.stack([sw, 1 - sw], axis=1).ravel().copy()
    ds = rng.normal(0, 0.5, n)
    O.update_fraction_pair(s, ds, w=1.0, abs_max=0.2, minval=0.0, maxval=1.0)
    maxval = min(1 - 0.0, 1.0 - 2 * 0.0); minval = max(0.0, maxval - 1)
    ref = np.array([swi + 1.0 * W._choose_increment(swi, di, 0.2, None, minval, maxval) for swi, di in zip(sw, ds)])
    assert np.array_equal(s[0::2], ref) and np.allclose(s[0::2] + s[1::2], 1.0, atol=1e-15)
