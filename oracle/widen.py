"""CPU restatements (numpy / plain Python) of the widened rows of the hot path.

TEST INFRASTRUCTURE ONLY — like the rest of oracle/: only tests/, __graft_entry__.smoke() and bench.py's CPU legs may
import it; nothing under jutul.jl_b200/ does. Every function cites the reference lines it follows. The reference
evaluates these functions on ForwardDiff Duals; `Dual` below is that arithmetic written out (value + partials, the rules
ForwardDiff uses for + - * / exp ^ abs and comparisons on the value), so the derivatives here are obtained the way the
reference obtains them and NOT by the closed forms the device kernels use.

Parity pins: interpolation against the known answers of test/utils.jl:154-246 (tests/test_oracle_widen.py); sort_symbols
restates Graphs.jl `topological_sort_by_dfs` (Graphs.jl is a dependency absent from /root/reference, compat "1.8.0" in
Project.toml — algorithm restated from the published source; no reference test pins an ordering => parity unpinned for
the order, pinned for the defining property: every variable after its dependencies); the :fvm assembly and the adjoint
alignment have no golden vectors in the reference => pinned by identities (finite differences, transposition).
"""
import math

import numpy as np

f64 = np.float64
i64 = np.int64


# ------------------------------------------------------------------ ForwardDiff.Dual written out
class Dual:
    __slots__ = ("v", "d")
    __array_ufunc__ = None             # ndarray (op) Dual defers to Dual.__r<op>__ instead of broadcasting over objects

    def __init__(self, v, d):
        self.v = v                     # float or ndarray (n,)
        self.d = d                     # ndarray (np,) or (np, n)

    @staticmethod
    def lift(x, like):
        return x if isinstance(x, Dual) else Dual(x, np.zeros_like(like.d))

    def __add__(self, o):
        o = Dual.lift(o, self)
        return Dual(self.v + o.v, self.d + o.d)
    __radd__ = __add__

    def __neg__(self):
        return Dual(-self.v, -self.d)

    def __sub__(self, o):
        o = Dual.lift(o, self)
        return Dual(self.v - o.v, self.d - o.d)

    def __rsub__(self, o):
        return Dual.lift(o, self) - self

    def __mul__(self, o):
        o = Dual.lift(o, self)
        return Dual(self.v * o.v, self.d * o.v + o.d * self.v)
    __rmul__ = __mul__

    def __truediv__(self, o):
        o = Dual.lift(o, self)
        q = self.v / o.v
        return Dual(q, (self.d - q * o.d) / o.v)

    def __rtruediv__(self, o):
        return Dual.lift(o, self) / self

    def __abs__(self):
        s = np.where(np.signbit(self.v), -1.0, 1.0)
        return Dual(np.abs(self.v), s * self.d)


def value(x):
    return x.v if isinstance(x, Dual) else x


def dexp(x):
    e = np.exp(value(x))
    return Dual(e, e * x.d) if isinstance(x, Dual) else e


def dpow(x, n):
    """x^n for a real exponent (ForwardDiff: n x^(n-1))."""
    if not isinstance(x, Dual):
        return x ** n
    with np.errstate(all="ignore"):
        return Dual(x.v ** n, np.where(x.v == 0.0, 0.0, n * x.v ** (n - 1.0)) * x.d)


def dclamp(x, lo, hi):
    """clamp(x, lo, hi): outside the interval the bound is returned as a constant (no partials)."""
    v = value(x)
    out_v = np.clip(v, lo, hi)
    inside = (v >= lo) & (v <= hi)
    return Dual(out_v, np.where(inside, 1.0, 0.0) * x.d) if isinstance(x, Dual) else out_v


# ------------------------------------------------------------------ src/interpolation.jl
def interpolation_constant_lookup(X, constant_dx=None):
    """interpolation_constant_lookup (interpolation.jl:50-67). None = `missing`. Returns (x0, dx, n) or None."""
    X = np.asarray(X, dtype=f64)
    dx = X[1] - X[0]
    if constant_dx is None:
        constant_dx = True
        for i in range(1, len(X)):
            b = X[i] - X[i - 1]
            constant_dx = constant_dx and abs(dx - b) <= math.sqrt(np.finfo(f64).eps) * max(abs(dx), abs(b))   # isapprox
    return (X[0], dx, len(X)) if constant_dx else None


def first_lower(tab, x, lookup=None):
    """first_lower (interpolation.jl:3-20), 1-based position."""
    if lookup is None:
        pos = int(np.searchsorted(tab, x, side="left")) + 1          # searchsortedfirst
        return min(max(pos - 1, 1), len(tab) - 1)
    x0, dx, n = lookup
    if x <= x0 + dx:
        return 1
    return min(int(math.floor((x - x0) / dx)) + 1, n - 1)


def linear_interp(X, F, x, lookup=None):
    """linear_interp / linear_interp_internal (interpolation.jl:29-40); x may be a Dual (scalar)."""
    ix = first_lower(X, float(value(x)), lookup) - 1
    X0, F0 = X[ix], F[ix]
    dFdX = (F[ix + 1] - F0) / (X[ix + 1] - X0)
    return F0 + dFdX * (x - X0)


def bilinear_interp(X, Y, F, x, y, lookup_x=None, lookup_y=None):
    """bilinear_interp (interpolation.jl:184-216); F[i, j] = f(X[i], Y[j])."""
    def interp_local(X0, F0, X1, F1, xx):
        return F0 + ((F1 - F0) / (X1 - X0)) * (xx - X0)
    ix = first_lower(X, float(value(x)), lookup_x) - 1
    iy = first_lower(Y, float(value(y)), lookup_y) - 1
    x1, x2, y1, y2 = X[ix], X[ix + 1], Y[iy], Y[iy + 1]
    Dy = y2 - y1
    F_upper = interp_local(x1, F[ix, iy + 1], x2, F[ix + 1, iy + 1], x)
    F_lower = interp_local(x1, F[ix, iy], x2, F[ix + 1, iy], x)
    w_lower = (y2 - y) / Dy
    w_upper = (y - y1) / Dy
    return w_lower * F_lower + w_upper * F_upper


class LinearInterpolant:
    """LinearInterpolant(X, F; constant_dx) (interpolation.jl:69-99)."""

    def __init__(self, X, F, constant_dx=None):
        X = np.array(X, dtype=f64); F = np.array(F, dtype=f64)
        if len(X) != len(F):
            raise ValueError("X and F values must have equal length.")
        if len(X) == 1:
            X = np.append(X, X[0] + 1.0); F = np.append(F, F[0])
        if np.any(np.diff(X) < 0):
            ix = np.argsort(X, kind="stable")
            X, F = X[ix], F[ix]
        self.X, self.F = X, F
        self.lookup = interpolation_constant_lookup(X, constant_dx)

    def __call__(self, x):
        return linear_interp(self.X, self.F, x, self.lookup)


def get_1d_interpolator(xs, ys, cap_endpoints=True, cap_start=None, cap_end=None, constant_dx=None):
    """get_1d_interpolator (interpolation.jl:118-154)."""
    cap_start = cap_endpoints if cap_start is None else cap_start
    cap_end = cap_endpoints if cap_end is None else cap_end
    xs = list(np.asarray(xs, dtype=f64)); ys = list(np.asarray(ys, dtype=f64))
    if len(xs) == 1:
        xs = [xs[0] - 1.0, xs[0], xs[0] + 1.0]; ys = [ys[0]] * 3
    elif cap_endpoints and (cap_start or cap_end):
        if cap_start:
            eps = xs[1] - xs[0]
            xs.insert(0, xs[0] - eps); ys.insert(0, ys[0])
        if cap_end:
            eps = xs[-1] - xs[-2]
            xs.append(xs[-1] + eps); ys.append(ys[-1])
    if len(xs) < 2:
        raise ValueError("xs values must have more than one entry.")
    return LinearInterpolant(xs, ys, constant_dx=constant_dx)


class BilinearInterpolant:
    """BilinearInterpolant (interpolation.jl:156-222)."""

    def __init__(self, xs, ys, fs, constant_dx=None, constant_dy=None):
        self.X = np.array(xs, dtype=f64); self.Y = np.array(ys, dtype=f64); self.F = np.array(fs, dtype=f64)
        if np.any(np.diff(self.X) < 0) or np.any(np.diff(self.Y) < 0):
            raise ValueError("xs and ys must be sorted.")
        if self.F.shape != (len(self.X), len(self.Y)):
            raise ValueError("f(x, y) must match lengths of xs and ys")
        self.lookup_x = interpolation_constant_lookup(self.X, constant_dx)
        self.lookup_y = interpolation_constant_lookup(self.Y, constant_dy)

    def __call__(self, x, y):
        return bilinear_interp(self.X, self.Y, self.F, x, y, self.lookup_x, self.lookup_y)


def get_2d_interpolator(xs, ys, fs, cap_endpoints=True, constant_dx=None, constant_dy=None):
    """get_2d_interpolator (interpolation.jl:224-285) with cap_x = cap_y = (cap_endpoints, cap_endpoints)."""
    xs = list(np.asarray(xs, dtype=f64)); ys = list(np.asarray(ys, dtype=f64)); fs = np.asarray(fs, dtype=f64)
    if cap_endpoints:
        nx, ny = fs.shape
        new = np.zeros((nx + 2, ny + 2))
        new[1:nx + 1, 1:ny + 1] = fs
        new[0, :] = new[1, :]
        xs.insert(0, xs[0] - (xs[1] - xs[0]))
        new[-1, :] = new[-2, :]
        xs.append(xs[-1] + (xs[-1] - xs[-2]))
        new[:, 0] = new[:, 1]
        ys.insert(0, ys[0] - (ys[1] - ys[0]))
        new[:, -1] = new[:, -2]
        ys.append(ys[-1] + (ys[-1] - ys[-2]))
        fs = new
    return BilinearInterpolant(xs, ys, fs, constant_dx=constant_dx, constant_dy=constant_dy)


# ------------------------------------------------------------------ src/variable_evaluation.jl
def sort_symbols(symbols, deps):
    """sort_symbols (variable_evaluation.jl:327-350): edges symbol -> dependency, reverse(topological_sort_by_dfs(graph)),
    i.e. the depth-first post-order over the vertices in declaration order with out-neighbours in ascending vertex order
    (Graphs.jl topological_sort_by_dfs). Returns 0-based indices."""
    n = len(symbols)
    adj = []
    for dep in deps:
        out = []
        for d in dep:
            pos = [i for i, s in enumerate(symbols) if s == d]
            if len(pos) != 1:
                raise KeyError(f"Symbol {d} must appear exactly once in secondary variables or parameters, found {len(pos)} entries.")
            out.append(pos[0])
        adj.append(sorted(set(out)))
    colour = [0] * n
    verts = []
    for v0 in range(n):
        if colour[v0]:
            continue
        S = [v0]
        colour[v0] = 1
        while S:
            u = S[-1]
            w = -1
            for nb in adj[u]:
                if colour[nb] == 1:
                    raise ValueError("The input graph contains at least one loop.")
                if colour[nb] == 0:
                    w = nb
                    break
            if w >= 0:
                colour[w] = 1
                S.append(w)
            else:
                colour[u] = 2
                verts.append(u)
                S.pop()
    return verts


def sort_secondary_variables(primary, secondary, parameters):
    """sort_secondary_variables! (variable_evaluation.jl:289-325). primary / parameters: lists of names; secondary: ordered
    dict name -> dependency list. Returns the sorted list of secondary names."""
    sec = list(secondary)
    for a, b, what in ((primary, sec, "primary and secondary variables"), (primary, parameters, "primary variables and parameters"),
                       (parameters, sec, "parameters and secondary variables")):
        isect = set(a) & set(b)
        if isect:
            raise ValueError(f"{sorted(isect)} found in both {what}.")
    nodes = list(primary) + list(parameters) + sec
    edges = [[] for _ in primary] + [[] for _ in parameters] + [list(secondary[k]) for k in sec]
    order = sort_symbols(nodes, edges)
    npar = len(primary) + len(parameters)
    return [nodes[i] for i in order if i >= npar]


def update_secondary_variables_state(state, definitions, order):
    """update_secondary_variables_state! (variable_evaluation.jl:110-148): `order` from sort_secondary_variables; `state` maps
    names to ndarray (parameters) or Dual (primaries, unit partial). definitions[name] = dict(kind, deps, c, table)."""
    like = next(v for v in state.values() if isinstance(v, Dual))
    for name in order:
        d = definitions[name]
        a = [state[k] for k in d.get("deps", [])]
        c = d["c"]
        k = d["kind"]
        if k == "const":
            out = Dual(np.full_like(like.v, c[0]), np.zeros_like(like.d))
        elif k == "affine":
            out = c[0]
            for j, aj in enumerate(a):
                out = out + c[1 + j] * aj
        elif k == "product":
            out = c[0]
            for aj in a:
                out = out * aj
        elif k == "quotient":
            out = c[0] * a[0] / a[1]
        elif k == "exp":
            out = c[0] * dexp(c[1] * (a[0] - c[2]))
        elif k == "power":
            out = c[0] * dpow(dclamp((a[0] - c[2]) / c[3], 0.0, 1.0), c[1])
        elif k == "table1d":
            I = d["table"]
            x = Dual.lift(a[0], like)
            vals = np.zeros_like(x.v); der = np.zeros_like(x.v)
            for i in range(x.v.shape[0]):
                r = I(Dual(x.v[i], np.array([1.0])))
                vals[i], der[i] = r.v, r.d[0]
            out = c[0] * Dual(vals, der[None, :] * x.d)
        elif k == "table2d":
            I = d["table"]
            x = Dual.lift(a[0], like); y = Dual.lift(a[1], like)
            vals = np.zeros_like(x.v); dx = np.zeros_like(x.v); dy = np.zeros_like(x.v)
            for i in range(x.v.shape[0]):
                r = I(Dual(x.v[i], np.array([1.0, 0.0])), Dual(y.v[i], np.array([0.0, 1.0])))
                vals[i], dx[i], dy[i] = r.v, r.d[0], r.d[1]
            out = c[0] * Dual(vals, dx[None, :] * x.d + dy[None, :] * y.d)
        else:
            raise ValueError(k)
        state[name] = Dual.lift(out, like) if not isinstance(out, Dual) else out
    return state


# ------------------------------------------------------------------ src/NFVM + src/conservation/fvm_assembly.jl
def nfvm_discretization_stencil(left, right, L, R=None):
    """discretization_stencil per face (NFVM/decomposition.jl:131-141, types.jl:38-42): unique!([left, right, mpfa cells...]),
    for the nonlinear discretisation unique!(vcat(stencil(ft_left), stencil(ft_right))). Returns (vpos, vars), 1-based."""
    nf = len(left)
    vpos = [1]
    vars_ = []
    for f in range(nf):
        cells = []

        def push(c):
            if c not in cells:
                cells.append(c)
        for D in ([L] if R is None else [L, R]):
            push(int(left[f])); push(int(right[f]))
            for k in range(int(D["ptr"][f]) - 1, int(D["ptr"][f + 1]) - 1):
                push(int(D["cell"][k]))
        vars_.extend(cells)
        vpos.append(len(vars_) + 1)
    return np.array(vpos, dtype=i64), np.array(vars_, dtype=i64)


def nfvm_evaluate_flux_dual(f, left, right, L, R, scheme, P):
    """evaluate_flux (NFVM/evaluation.jl:1-88) for face f on a pressure accessor P(cell) returning floats or Duals."""
    def tpfa_flux(p_l, p_r, D):
        return D["T_left"][f] * p_l + D["T_right"][f] * p_r

    def compute_r(D):
        q = 0.0
        for k in range(int(D["ptr"][f]) - 1, int(D["ptr"][f + 1]) - 1):
            q = q + P(int(D["cell"][k])) * D["T"][k]
        return q

    def half(D, sgn=None):
        q = tpfa_flux(p_l, p_r, D)
        r = compute_r(D)
        q = q + r
        return (q, r) if sgn is None else (sgn * q, sgn * r)
    p_l, p_r = P(int(left[f])), P(int(right[f]))
    if scheme == "linear":
        return tpfa_flux(p_l, p_r, L) + compute_r(L)
    q_l, r_l = half(L)
    q_r, r_r = half(R, -1)
    if scheme == "ntpfa":
        r_lw, r_rw = r_l, r_r
    else:
        r_lw, r_rw = abs(r_l), abs(r_r)
    r_total = r_lw + r_rw
    if abs(value(r_total)) < 1e-10:
        mu_l = mu_r = 0.5
    else:
        mu_l = r_rw / r_total
        mu_r = r_lw / r_total
    return mu_l * q_l - mu_r * q_r


def fvm_declare_pattern(nc, left, right, vpos, vars_):
    """declare_pattern for the :fvm storage (fvm_assembly.jl:55-89): (I, J), 1-based, before unique / sparse()."""
    I = list(range(1, nc + 1)); J = list(range(1, nc + 1))
    for f in range(len(left)):
        lc, rc = int(left[f]), int(right[f])
        for s in range(vpos[f] - 1, vpos[f + 1] - 1):
            c = int(vars_[s])
            I += [lc, rc, c, c]; J += [c, c, lc, rc]
    return np.array(I, dtype=i64), np.array(J, dtype=i64)


def fvm_align(left, right, vpos, vars_, rowptr, colidx):
    """align_to_jacobian! (fvm_assembly.jl:98-165) for ne = np = 1 on a CSR Jacobian: left / right nz index per slot (1-based)."""
    def find(row, col):
        for k in range(rowptr[row - 1] - 1, rowptr[row] - 1):
            if colidx[k] == col:
                return k + 1
        raise RuntimeError("Jacobian alignment failed. Giving up.")
    lp = np.zeros(len(vars_), dtype=i64); rp = np.zeros(len(vars_), dtype=i64)
    for f in range(len(left)):
        for s in range(vpos[f] - 1, vpos[f + 1] - 1):
            lp[s] = find(int(left[f]), int(vars_[s])); rp[s] = find(int(right[f]), int(vars_[s]))
    return lp, rp


def fvm_assemble_nfvm(nc, left, right, L, R, scheme, p, vpos, vars_, lpos, rpos, diag_pos, nnz, acc=None, dacc=None):
    """update_equation! + update_linearized_system_equation! (fvm_assembly.jl:178-283) for the pressure law on the NFVM flux:
    the face flux is evaluated once per stencil slot with that cell carrying the partial (fvm_update_face_fluxes_inner!),
    then fvm_face_assembly! scatters in face order. Returns (nz, r, q)."""
    nz = np.zeros(nnz); r = np.zeros(nc); q = np.zeros(len(left))
    for c in range(nc):
        r[c] = 0.0 if acc is None else acc[c]
        if dacc is not None:
            nz[diag_pos[c] - 1] = dacc[c]
    for f in range(len(left)):
        lc, rc = int(left[f]) - 1, int(right[f]) - 1
        start, stop = vpos[f] - 1, vpos[f + 1] - 1
        entries = []
        for s in range(start, stop):
            var = int(vars_[s])
            P = lambda c, var=var: Dual(p[c - 1], np.array([1.0 if c == var else 0.0]))
            entries.append(nfvm_evaluate_flux_dual(f, left, right, L, R, scheme, P))
        qval = entries[0].v
        q[f] = qval
        r[lc] += qval
        r[rc] -= qval
        for s, e in zip(range(start, stop), entries):
            nz[lpos[s] - 1] += e.d[0]
            nz[rpos[s] - 1] -= e.d[0]
    return nz, r, q


# ------------------------------------------------------------------ adjoint layouts (src/equations.jl:101-108,152-161)
def find_jac_position_block(rowptr, colidx, row, col, eq, partial, N, adjoint=False):
    """find_jac_position for BlockMajorLayout on a CSR Jacobian: 1-based index into the flat nzval. With as_adjoint the
    block looked up is (col, row) and the in-block index is N (eq - 1) + partial."""
    a, b = (col, row) if adjoint else (row, col)
    for k in range(rowptr[a - 1] - 1, rowptr[a] - 1):
        if colidx[k] == b:
            pos = k + 1
            base = (pos - 1) * N * N
            return base + (N * (eq - 1) + partial if adjoint else N * (partial - 1) + eq)
    raise RuntimeError("Jacobian alignment failed. Giving up.")


# ------------------------------------------------------------------ two-phase law on property Duals
def assemble_2ph_props(hf, diag_pos, hf_pos, Tf, gdz, p, props, M0, dt, nnzb, src_cells=None, src_vals=None):
    """update_accumulation! + update_half_face_flux_tpfa! + fill_conservation_eq! (src/conservation/conservation.jl:558-626,
    373-430) for the two-phase law whose properties are secondary variables held as Duals in the state: props[name] is a
    (3, nc) array — value, d/dp, d/dSw of the cell's own primaries — for MassW, MassO, DensityW, DensityO, MobilityW, MobilityO.
    Local-perspective AD (src/ad/local_ad.jl:54-79): inside the loop of cell c only the entries of c carry partials, the
    neighbour's are plain values. The flux follows two_point_potential_drop / upw_flux (src/conservation/flux.jl:335-435).
    Same summation order as the reference: accumulation first, then the half-faces in conn_pos order. 1-based tables."""
    nc = p.shape[0]
    nz = np.zeros(nnzb * 4); r = np.zeros(2 * nc)
    src = np.zeros((nc, 2))
    if src_cells is not None:
        for c, v in zip(src_cells, np.asarray(src_vals).reshape(-1, 2)):
            src[int(c) - 1] += v
    names = (("MassW", "DensityW", "MobilityW"), ("MassO", "DensityO", "MobilityO"))

    def dual(a, c):
        return Dual(a[0, c], np.array([a[1, c], a[2, c]]))
    fp, cells, faces, sign = hf["face_pos"], hf["cells"], hf["faces"], hf["face_sign"]
    for c in range(nc):
        ps = Dual(p[c], np.array([1.0, 0.0]))
        for a in range(2):
            mass, rho, mob = (props[k] for k in names[a])
            acc = (dual(mass, c) - M0[2 * c + a]) / dt + src[c, a]
            rv, dv = acc.v, acc.d.copy()
            for i in range(fp[c] - 1, fp[c + 1] - 1):
                o = int(cells[i]) - 1
                f = int(faces[i]) - 1
                sg = sign[i] * gdz[f]
                rho_avg = 0.5 * (dual(rho, c) + rho[0, o])
                theta = ps - p[o] + sg * rho_avg
                q = Tf[f] * theta
                F = (dual(mob, c) * q) if q.v > 0 else (mob[0, o] * q)
                rv += F.v
                dv += F.d
                pos = int(hf_pos[i])                       # block (other, self): -dF/dx_self
                for d in range(2):
                    nz[(pos - 1) * 4 + 2 * d + a] = -F.d[d]
            r[2 * c + a] = rv
            for d in range(2):
                nz[(int(diag_pos[c]) - 1) * 4 + 2 * d + a] = dv[d]
    return nz, r


# ------------------------------------------------------------------ unit_sum_update! for nf > 2 (src/variables/utils.jl:393-526)
MINIMUM_SAT_RELAX = 1e-3


def _choose_increment(v, dv, abs_change=None, rel_change=None, minval=None, maxval=None, scale=None):
    """choose_increment (src/variables/utils.jl:149-174)"""
    if scale is not None:
        dv = dv * scale
    if abs_change is not None:
        dv = np.sign(dv) * min(abs(dv), abs_change)
    if rel_change is not None:
        dv = np.sign(dv) * min(abs(dv), rel_change * abs(v))
    if minval is not None:
        dv = max(dv, minval - v)
    if maxval is not None:
        dv = min(dv, maxval - v)
    return dv


def _pick_relaxation(w, dv, dv0):
    with np.errstate(all="ignore"):
        r = np.float64(dv) / np.float64(dv0)
    if dv0 != 0:
        w = min(w, r)
    return min(w, MINIMUM_SAT_RELAX)


def _unit_update_magnitude_local(s, dx, nf, minval, maxval, abs_max):
    dlast0 = 0.0
    for i in range(nf - 1):
        dv = _choose_increment(s[i], dx[i], abs_max, None, minval, maxval)
        s[i] += dv
        dlast0 -= dv
    dlast = _choose_increment(s[nf - 1], dlast0, abs_max, None, minval, maxval)
    s[nf - 1] += dlast
    if dlast != dlast0:
        t = 0.0
        for i in range(nf):
            t += s[i]
        for i in range(nf):
            s[i] = min(max(s[i], minval), maxval) / t


def unit_sum_update(s, dx, nf, w0=1.0, abs_max=None, minval=0.0, maxval=1.0, preserve_direction=True):
    """unit_sum_update! for nf > 2: s is (n, nf), dx is (n, nf - 1). In place."""
    maxval = maxval - nf * minval
    for c in range(s.shape[0]):
        sc, dc = s[c], dx[c]
        if not preserve_direction:
            _unit_update_magnitude_local(sc, dc, nf, minval, maxval, abs_max)
            continue
        w = 1.0; dlast0 = 0.0
        for i in range(nf - 1):
            dv = _choose_increment(sc[i], dc[i], abs_max, None, minval, maxval)
            dlast0 -= dc[i]
            w = _pick_relaxation(w, dv, dc[i])
        dlast = _choose_increment(sc[nf - 1], dlast0, abs_max, None, minval, maxval)
        w = w0 * _pick_relaxation(w, dlast, dlast0)
        if w <= MINIMUM_SAT_RELAX:
            _unit_update_magnitude_local(sc, dc, nf, minval, maxval, abs_max)
        else:
            for i in range(nf - 1):
                sc[i] += w * dc[i]
            sc[nf - 1] += w * dlast0
    return s
