"""ctypes front-end of the CPU oracle (oracle/oracle.cpp).

TEST INFRASTRUCTURE ONLY — see the header of oracle.cpp. Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package. All index arrays are 1-based int64 like the reference's.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liborc.so")

i64 = np.int64
f64 = np.float64
_pi = C.POINTER(C.c_int64)
_pd = C.POINTER(C.c_double)


def build(force=False):
    src = os.path.join(_HERE, "oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.orc_cart_num_faces.restype = C.c_int64
        _lib.orc_csr_from_coo.restype = C.c_int64
        _lib.orc_block_nz_index.restype = C.c_int64
        _lib.orc_index_equation_major.restype = C.c_int64
        _lib.orc_index_entity_major.restype = C.c_int64
        _lib.orc_ilu0_create.restype = C.c_void_p
        _lib.orc_ilu0_nnz_l.restype = C.c_int64
        _lib.orc_ilu0_nnz_u.restype = C.c_int64
        _lib.orc_ilu0_factor.restype = C.c_int
        _lib.orc_bicgstab.restype = C.c_int
        _lib.orc_num_threads.restype = C.c_int
    return _lib


def _I(a):
    return None if a is None else a.ctypes.data_as(_pi)


def _D(a):
    return None if a is None else a.ctypes.data_as(_pd)


def _ci(x):
    return C.c_int64(int(x))


def _cd(x):
    return C.c_double(float(x))


def _ai(a):
    return np.ascontiguousarray(a, dtype=i64)


def _ad(a):
    return np.ascontiguousarray(a, dtype=f64)


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(n):
    lib().orc_set_num_threads(C.c_int(n))


# ------------------------------------------------------------------ mesh
def cart_neighbors(nx, ny=1, nz=1):
    """2 x nf neighbourship (1-based) in MRST order; returned as (nf, 2) array
    whose memory equals Julia's column-major 2 x nf."""
    nf = lib().orc_cart_num_faces(_ci(nx), _ci(ny), _ci(nz))
    N = np.zeros((nf, 2), dtype=i64)
    lib().orc_cart_neighbors(_ci(nx), _ci(ny), _ci(nz), _I(N))
    return N


def cart_geometry(nx, ny, nz, dx, dy, dz):
    nc = nx * ny * nz
    nf = lib().orc_cart_num_faces(_ci(nx), _ci(ny), _ci(nz))
    cc = np.zeros((nc, 3)); vol = np.zeros(nc); fc = np.zeros((nf, 3)); fn = np.zeros((nf, 3)); fa = np.zeros(nf)
    lib().orc_cart_geometry(_ci(nx), _ci(ny), _ci(nz), _cd(dx), _cd(dy), _cd(dz), _D(cc), _D(vol), _D(fc), _D(fn), _D(fa))
    return dict(cell_centroids=cc, volumes=vol, face_centroids=fc, normals=fn, areas=fa)


def half_face_map(N, nc):
    N = _ai(N)
    nf = N.shape[0]
    nhf = 2 * nf
    faces = np.zeros(nhf, dtype=i64); face_pos = np.zeros(nc + 1, dtype=i64)
    cells = np.zeros(nhf, dtype=i64); signs = np.zeros(nhf, dtype=i64)
    lib().orc_half_face_map(_I(N), _ci(nf), _ci(nc), _I(faces), _I(face_pos), _I(cells), _I(signs))
    return dict(cells=cells, faces=faces, face_pos=face_pos, face_sign=signs)


def half_face_trans(geo, perm, hf):
    nc = geo["volumes"].shape[0]
    perm = _ad(np.broadcast_to(perm, (nc,)))
    T = np.zeros(hf["faces"].shape[0])
    lib().orc_half_face_trans(_ci(nc), _D(geo["cell_centroids"]), _D(geo["face_centroids"]), _D(geo["normals"]),
                              _D(geo["areas"]), _D(perm), _I(hf["faces"]), _I(hf["face_pos"]), _I(hf["face_sign"]), _D(T))
    return T


def face_trans(T_hf, faces, nf):
    T = np.zeros(nf)
    lib().orc_face_trans(_ci(nf), _ci(T_hf.shape[0]), _D(_ad(T_hf)), _I(faces), _D(T))
    return T


def face_gdz(N, z, g=9.80665):
    N = _ai(N)
    gdz = np.zeros(N.shape[0])
    lib().orc_face_gdz(_I(N), _ci(N.shape[0]), _D(_ad(z)), _cd(g), _D(gdz))
    return gdz


# ------------------------------------------------------------------ pattern
def csr_from_coo(I, J, n):
    I = _ai(I); J = _ai(J)
    rowptr = np.zeros(n + 1, dtype=i64)
    nnz = lib().orc_csr_from_coo(_I(I), _I(J), _ci(I.shape[0]), _ci(n), _I(rowptr), None)
    colidx = np.zeros(nnz, dtype=i64)
    lib().orc_csr_from_coo(_I(I), _I(J), _ci(I.shape[0]), _ci(n), _I(rowptr), _I(colidx))
    return rowptr, colidx


def tpfa_pattern(hf):
    nc = hf["face_pos"].shape[0] - 1
    nhf = hf["faces"].shape[0]
    I = np.zeros(nhf + nc, dtype=i64); J = np.zeros(nhf + nc, dtype=i64)
    lib().orc_tpfa_pattern(_ci(nc), _ci(nhf), _I(hf["face_pos"]), _I(hf["cells"]), _I(I), _I(J))
    return I, J


def tpfa_alignment(hf, rowptr, colidx):
    nc = hf["face_pos"].shape[0] - 1
    diag_pos = np.zeros(nc, dtype=i64); hf_pos = np.zeros(hf["faces"].shape[0], dtype=i64)
    lib().orc_tpfa_alignment(_ci(nc), _I(hf["face_pos"]), _I(hf["cells"]), _I(rowptr), _I(colidx), _I(diag_pos), _I(hf_pos))
    return diag_pos, hf_pos


def block_nz_index(pos, N, eq, partial):
    return lib().orc_block_nz_index(_ci(pos), _ci(N), _ci(eq), _ci(partial))


def index_equation_major(outer, inner, n_outer):
    return lib().orc_index_equation_major(_ci(outer), _ci(inner), _ci(n_outer))


def index_entity_major(outer, inner, n_inner):
    return lib().orc_index_entity_major(_ci(outer), _ci(inner), _ci(n_inner))


# ------------------------------------------------------------------ assembly
def params_array(rho0=(1000.0, 700.0), c=(4.5e-10, 1e-9), mu=(1e-3, 5e-3), p0=1e5):
    return np.array([rho0[0], rho0[1], c[0], c[1], mu[0], mu[1], p0], dtype=f64)


def assemble_2ph(hf, diag_pos, hf_pos, Tf, gdz, pv, params, p, sw, M0, dt, nnzb, src_cells=None, src_vals=None):
    nc = p.shape[0]
    nz = np.zeros(nnzb * 4); r = np.zeros(2 * nc)
    nsrc = 0 if src_cells is None else len(src_cells)
    sc = None if nsrc == 0 else _ai(src_cells)
    sv = None if nsrc == 0 else _ad(src_vals)
    lib().orc_assemble_2ph(_ci(nc), _I(hf["face_pos"]), _I(hf["faces"]), _I(hf["cells"]), _I(hf["face_sign"]),
                           _I(diag_pos), _I(hf_pos), _D(_ad(Tf)), _D(_ad(gdz)), _D(_ad(pv)), _D(params),
                           _D(_ad(p)), _D(_ad(sw)), _D(_ad(M0)), _cd(dt), _ci(nsrc), _I(sc), _D(sv), _D(nz), _D(r))
    return nz, r


def mass_2ph(pv, params, p, sw):
    nc = p.shape[0]
    M = np.zeros(2 * nc)
    lib().orc_mass_2ph(_ci(nc), _D(_ad(pv)), _D(params), _D(_ad(p)), _D(_ad(sw)), _D(M))
    return M


def residual_2ph(hf, Tf, gdz, pv, params, p, sw, M0, dt):
    nc = p.shape[0]
    r = np.zeros(2 * nc)
    lib().orc_residual_2ph(_ci(nc), _I(hf["face_pos"]), _I(hf["faces"]), _I(hf["cells"]), _I(hf["face_sign"]),
                           _D(_ad(Tf)), _D(_ad(gdz)), _D(_ad(pv)), _D(params), _D(_ad(p)), _D(_ad(sw)), _D(_ad(M0)),
                           _cd(dt), _D(r))
    return r


def heat_pattern(nx, ny):
    nc = nx * ny
    I = np.zeros(5 * nc, dtype=i64); J = np.zeros(5 * nc, dtype=i64)
    lib().orc_heat_pattern(_ci(nx), _ci(ny), _I(I), _I(J))
    return I, J


def assemble_heat(nx, ny, hx, hy, dt, T, T0, rowptr, colidx):
    nz = np.zeros(colidx.shape[0]); r = np.zeros(nx * ny)
    lib().orc_assemble_heat(_ci(nx), _ci(ny), _cd(hx), _cd(hy), _cd(dt), _D(_ad(T)), _D(_ad(T0)), _I(rowptr), _I(colidx),
                            _D(nz), _D(r))
    return nz, r


def assemble_poisson(hf, rowptr, colidx, K, U, U0, time_dependent, dt, src_cells=None, src_vals=None):
    nc = U.shape[0]
    nz = np.zeros(colidx.shape[0]); r = np.zeros(nc)
    nsrc = 0 if src_cells is None else len(src_cells)
    sc = None if nsrc == 0 else _ai(src_cells)
    sv = None if nsrc == 0 else _ad(src_vals)
    lib().orc_assemble_poisson(_ci(nc), _I(hf["face_pos"]), _I(hf["faces"]), _I(hf["cells"]), _I(rowptr), _I(colidx),
                               _D(_ad(K)), _D(_ad(U)), _D(_ad(U0)), C.c_int(int(time_dependent)), _cd(dt), _ci(nsrc),
                               _I(sc), _D(sv), _D(nz), _D(r))
    return nz, r


# ------------------------------------------------------------------ linear algebra
def spmv(n, bs, rowptr, colidx, nz, x, alpha=1.0, beta=0.0, y=None):
    if y is None:
        y = np.zeros(n * bs)
    lib().orc_spmv(_ci(n), C.c_int(bs), _I(rowptr), _I(colidx), _D(nz), _cd(alpha), _D(_ad(x)), _cd(beta), _D(y))
    return y


class ILU0:
    """ilu0_csr(A[, partition]) / ilu0_csr! / ldiv! (src/StaticCSR/ilu0.jl, par_ilu0.jl)."""

    def __init__(self, n, bs, rowptr, colidx, partition=None):
        self.n, self.bs = n, bs
        self._keep = (rowptr, colidx)
        part = None if partition is None else _ai(partition)
        self.h = C.c_void_p(lib().orc_ilu0_create(_ci(n), C.c_int(bs), _I(rowptr), _I(colidx), _I(part)))

    def factor(self, nz):
        return lib().orc_ilu0_factor(self.h, _D(_ad(nz)))

    def set_level_schedule(self, on=True, max_levels=64):
        """Run the triangular solves level by level (rows of a dependency level concurrently). Bitwise the same result as the
        sequential sweep; only worth it for orderings with few levels. Returns the number of levels (0: refused)."""
        f = lib().orc_ilu0_set_level_schedule
        f.restype = C.c_int64
        return int(f(self.h, C.c_int(1 if on else 0), _ci(max_levels)))

    def solve(self, b, x=None):
        if x is None:
            x = np.zeros_like(b)
        lib().orc_ilu0_solve(self.h, _D(_ad(b)), _D(x))
        return x

    def get(self):
        nl = lib().orc_ilu0_nnz_l(self.h); nu = lib().orc_ilu0_nnz_u(self.h)
        b2 = self.bs * self.bs
        L = np.zeros(nl * b2); U = np.zeros(nu * b2); D = np.zeros(self.n * b2)
        Lptr = np.zeros(self.n + 1, dtype=i64); Uptr = np.zeros(self.n + 1, dtype=i64)
        Lcol = np.zeros(nl, dtype=i64); Ucol = np.zeros(nu, dtype=i64)
        lib().orc_ilu0_get(self.h, _D(L), _D(U), _D(D), _I(Lptr), _I(Lcol), _I(Uptr), _I(Ucol))
        return dict(L=L, U=U, D=D, Lptr=Lptr, Lcol=Lcol, Uptr=Uptr, Ucol=Ucol)

    def __del__(self):
        try:
            lib().orc_ilu0_destroy(self.h)
        except Exception:
            pass


def bicgstab(n, bs, rowptr, colidx, nz, b, ilu=None, side="right", rtol=1e-3, atol=1e-12, itmax=100, min_it=1):
    x = np.zeros(n * bs)
    hist = np.zeros(itmax + 2)
    iters = C.c_int64(0)
    s = {"right": 0, "left": 1, "none": -1}[side]
    h = ilu.h if ilu is not None else None
    st = lib().orc_bicgstab(_ci(n), C.c_int(bs), _I(rowptr), _I(colidx), _D(_ad(nz)), h, C.c_int(s), _D(_ad(b)), _D(x),
                            _cd(rtol), _cd(atol), _ci(itmax), _ci(min_it), C.byref(iters), _D(hist), _ci(hist.shape[0]))
    return x, st, iters.value, hist[: iters.value + 1].copy()


def gmres(n, bs, rowptr, colidx, nz, b, ilu=None, side="right", rtol=1e-3, atol=1e-12, itmax=100, memory=20, restart=False, flexible=False):
    x = np.zeros(n * bs)
    hist = np.zeros(itmax + 2)
    iters = C.c_int64(0)
    s = {"right": 0, "left": 1, "none": -1}[side]
    h = ilu.h if ilu is not None else None
    lib().orc_gmres.restype = C.c_int
    st = lib().orc_gmres(_ci(n), C.c_int(bs), _I(rowptr), _I(colidx), _D(_ad(nz)), h, C.c_int(s), _D(_ad(b)), _D(x), _cd(rtol), _cd(atol),
                         _ci(itmax), _ci(memory), C.c_int(int(restart)), C.c_int(int(flexible)), C.byref(iters), _D(hist), _ci(hist.shape[0]))
    return x, st, iters.value, hist[: iters.value + 1].copy()


def nfvm_evaluate_flux(left, right, L, R, p, nph=1, ph=1, scheme="linear"):
    sc = {"linear": 0, "ntpfa": 1, "nmpfa": 2}[scheme]
    left = _ai(left); right = _ai(right)
    nf = left.shape[0]
    q = np.zeros(nf)

    def arrs(d):
        if d is None:
            return [None] * 5
        return [_ad(d["T_left"]), _ad(d["T_right"]), _ai(d["ptr"]), _ai(d["cell"]), _ad(d["T"])]
    a, b = arrs(L), arrs(R)
    lib().orc_nfvm_evaluate_flux(_ci(nf), C.c_int(sc), _I(left), _I(right), _D(a[0]), _D(a[1]), _I(a[2]), _I(a[3]), _D(a[4]),
                                 _D(b[0]), _D(b[1]), _I(b[2]), _I(b[3]), _D(b[4]), _D(_ad(p)), _ci(nph), _ci(ph), _D(q))
    return q


# ------------------------------------------------------------------ Newton glue
_NAN = float("nan")


def _opt(x):
    return _NAN if x is None else float(x)


def update_scalar(v, dx, w=1.0, abs_max=None, rel_max=None, minv=None, maxv=None, scale=None, dx_stride=1):
    lib().orc_update_scalar(_ci(v.shape[0]), _D(v), _D(dx), _ci(dx_stride), _cd(w), _cd(_opt(abs_max)), _cd(_opt(rel_max)),
                            _cd(_opt(minv)), _cd(_opt(maxv)), _cd(_opt(scale)))
    return v


def update_fraction_pair(s, dx, w=1.0, abs_max=None, minval=0.0, maxval=1.0, dx_stride=1):
    lib().orc_update_fraction_pair(_ci(s.shape[0] // 2), _D(s), _D(dx), _ci(dx_stride), _cd(w), _cd(_opt(abs_max)),
                                   _cd(minval), _cd(maxval))
    return s


def increment_norm(dx, stride=1, n=None):
    if n is None:
        n = dx.shape[0] // stride
    a = C.c_double(0); b = C.c_double(0)
    lib().orc_increment_norm(_ci(n), _D(dx), _ci(stride), C.byref(a), C.byref(b))
    return a.value, b.value


def maxabs_rows(r, bs):
    out = np.zeros(bs)
    lib().orc_maxabs_rows(_ci(r.shape[0] // bs), C.c_int(bs), _D(_ad(r)), _D(out))
    return out


def scale_dt(nz, r, bs, dt):
    lib().orc_scale_dt(_ci(nz.shape[0] // (bs * bs)), _ci(r.shape[0] // bs), C.c_int(bs), _cd(dt), _D(nz), _D(r))


def scale_diagonal(n, bs, rowptr, colidx, nz, r):
    lib().orc_scale_diagonal(_ci(n), C.c_int(bs), _I(rowptr), _I(colidx), _D(nz), _D(r))


def process_partition(N, nc, partition, weights=None):
    N = _ai(N); partition = _ai(partition)
    out = np.zeros(nc, dtype=i64)
    w = None if weights is None else _ad(weights)
    lib().orc_process_partition(_I(N), _ci(N.shape[0]), _ci(nc), _I(partition), _D(w), _I(out))
    return out


def partition_linear(m, n):
    p = np.zeros(n, dtype=i64)
    lib().orc_partition_linear(_ci(m), _ci(n), _I(p))
    return p


# ------------------------------------------------------------------ generic-cache equations (src/ad/generic.jl)
def generic_fill(ne, npart, vpos, dpos, positions, entries, nz, r, r_offset=0):
    """fill_equation_entries_impl! (src/ad/generic.jl:61-96), plain loops. vpos (nu+1), dpos (nu or None) and positions
    ((ne*np) x n_slots, column-major flat, 0 = not aligned) are 1-based like the cache's; entries is the memory image of
    Matrix{Dual}(ne, n_slots): shape (n_slots, ne, 1+np). nz (flat nzval) and r (ne x nu, equation fastest) are updated
    in place: Jacobian entries are SET (update_jacobian_inner!, src/ad/ad.jl:74-76)."""
    nu = len(vpos) - 1
    E = np.asarray(entries, dtype=f64).reshape(-1, ne, 1 + npart)
    P = np.asarray(positions, dtype=i64).reshape(-1, ne * npart)          # [slot][(e-1)*np + d-1]
    for i in range(nu):
        rs = (dpos[i] if dpos is not None else vpos[i])
        for jno, j in enumerate(range(vpos[i], vpos[i + 1])):
            fill_residual = (j == rs) if dpos is not None else (jno == 0)
            for e in range(ne):
                a = E[j - 1, e]
                if fill_residual:
                    r[r_offset + e + ne * i] = a[0]
                for d in range(npart):
                    q = P[j - 1, e * npart + d]
                    if q > 0:
                        nz[q - 1] = a[1 + d]
    return nz, r


def generic_cache_heat(nx, ny, hx, hy, dt, T, T0, rowptr, colidx):
    """What the reference holds for SimpleHeatEquation in its GenericAutoDiffCache after update_equation_in_entity!
    (src/applications/test_systems/heat_2d/heat_2d.jl:7-49) evaluated with the local perspective of every stencil cell
    (src/ad/local_ad.jl:54-79): per cell the unique stencil cells (aliases of a tiny periodic grid merged, as the
    sparsity tracing does), per slot the value of the equation and its partial w.r.t. T of that slot's cell, and the
    aligned positions in the CSR Jacobian (find_sparse_position, src/equations.jl:140-188). Returns
    (vpos, dpos, positions, entries) in the layout of generic_fill."""
    nc = nx * ny
    T = np.asarray(T, dtype=f64); T0 = np.asarray(T0, dtype=f64)
    vpos = [1]; dpos = []; pos = []; ent = []
    for c in range(nc):
        i, j = c % nx, c // nx
        L = (i - 1) % nx + j * nx; R = (i + 1) % nx + j * nx
        U = i + ((j - 1) % ny) * nx; D = i + ((j + 1) % ny) * nx
        d2x = (T[L] - 2 * T[c] + T[R]) / hx ** 2
        d2y = (T[U] - 2 * T[c] + T[D]) / hy ** 2
        val = (T[c] - T0[c]) / dt - (d2x + d2y)
        part = {}
        for cell, w in ((c, 1.0 / dt + 2.0 / hx ** 2 + 2.0 / hy ** 2), (L, -1.0 / hx ** 2), (R, -1.0 / hx ** 2), (U, -1.0 / hy ** 2),
                        (D, -1.0 / hy ** 2)):
            part[cell] = part.get(cell, 0.0) + w
        for cell in sorted(part):
            if cell == c:
                dpos.append(len(ent) + 1)
            row = colidx[rowptr[c] - 1: rowptr[c + 1] - 1]
            k = int(np.nonzero(row == cell + 1)[0][0])
            pos.append(rowptr[c] + k)          # 1-based flat index (scalar Jacobian: block index == nzval index)
            ent.append((val, part[cell]))
        vpos.append(len(ent) + 1)
    return (np.array(vpos, dtype=i64), np.array(dpos, dtype=i64), np.array(pos, dtype=i64).reshape(-1, 1),
            np.array(ent, dtype=f64).reshape(-1, 1, 2))


# ------------------------------------------------------------------ diagonal preconditioners (src/linsolve/precond)
def _block_rows(n, bs, rowptr, colidx, nz):
    b2 = bs * bs
    V = np.asarray(nz, dtype=f64).reshape(-1, b2)
    for i in range(n):
        yield i, range(rowptr[i] - 1, rowptr[i + 1] - 1), V


def jacobi_factor(n, bs, rowptr, colidx, nz, w=2.0 / 3.0):
    """JacobiPreconditioner (src/linsolve/precond/jacobi.jl:15-18): D_i = w * inv(A_ii), blocks column-major like nzval."""
    D = np.zeros((n, bs * bs))
    for i, ks, V in _block_rows(n, bs, rowptr, colidx, nz):
        for k in ks:
            if colidx[k] == i + 1:
                Aii = V[k].reshape(bs, bs).T                     # column-major block -> [row, col]
                D[i] = (w * np.linalg.inv(Aii)).T.ravel()
    return D.ravel()


def spai0_factor(n, bs, rowptr, colidx, nz):
    """SPAI0Preconditioner on a StaticSparsityMatrixCSR (src/linsolve/precond/spai.jl:42-62):
    D_row = inv(sum over the row's entries of the squared Frobenius norm) * A_ii."""
    D = np.zeros((n, bs * bs))
    for i, ks, V in _block_rows(n, bs, rowptr, colidx, nz):
        norm_sum = 0.0
        Aii = np.zeros(bs * bs)
        for k in ks:
            for v in V[k]:
                norm_sum += v * v
            if colidx[k] == i + 1:
                Aii = V[k]
        D[i] = (1.0 / norm_sum) * Aii
    return D.ravel()


def diagonal_apply(D, y, bs):
    """apply!(x, ::DiagonalPreconditioner, y) (src/linsolve/precond/diagonal.jl:31-49): x_i = D_i * y_i per block."""
    n = y.shape[0] // bs
    Dm = np.asarray(D, dtype=f64).reshape(n, bs, bs).transpose(0, 2, 1)      # [row, col]
    return np.einsum("nij,nj->ni", Dm, np.asarray(y, dtype=f64).reshape(n, bs)).ravel()


# ------------------------------------------------------------------ MultiModel Schur complement (src/linsolve/multimodel.jl)
# [B C; D E] [x; y] = [a; b] with the eliminated groups (wells, ...) in E: the Krylov solver sees S = B - sum_i C_i E_i^{-1} D_i.
# B is any object with a @-product (scipy sparse / ndarray); C_i, D_i, E_i are small (dense or sparse) blocks.
def _esolve(E, v):
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    return spla.spsolve(sp.csc_matrix(E), v) if sp.issparse(E) else np.linalg.solve(np.asarray(E), v)


def schur_prepare(a, C, E, b):
    """prepare_linear_solve! (multimodel.jl:17-33): a -= sum_i C_i (E_i \\ b_i), in place."""
    for Ci, Ei, bi in zip(C, E, b):
        a -= Ci @ _esolve(Ei, bi)
    return a


def schur_mul(B, C, D, E, x, alpha=1.0, beta=0.0, res=None):
    """schur_mul! (multimodel.jl:150-160): res <- beta res + alpha (B x - sum_i C_i (E_i \\ (D_i x)))."""
    out = alpha * (B @ x) + (0.0 if res is None or beta == 0.0 else beta * res)
    for Ci, Di, Ei in zip(C, D, E):
        out = out - alpha * (Ci @ _esolve(Ei, Di @ x))
    return out


def schur_dx_update(D, E, b, dx_from_solver):
    """update_dx_from_vector! / schur_dx_update! (multimodel.jl:97-137): x = -dx_from_solver; y_i = E_i \\ (D_i dx - b_i)."""
    x = -np.asarray(dx_from_solver, dtype=f64)
    y = [_esolve(Ei, Di @ dx_from_solver - bi) for Di, Ei, bi in zip(D, E, b)]
    return x, y


# ------------------------------------------------------------------ load-balanced intervals (src/partitioning.jl:310-336)
def load_balanced_endpoint(block_index, nvals, nblocks):
    """Endpoint of interval `block_index` when nvals is cut into nblocks, the first (nvals mod nblocks) blocks one wider."""
    width, remainder = divmod(nvals, nblocks)
    return min(min(block_index, remainder) + width * block_index, nvals)


def load_balanced_interval(b, n, m):
    """1-based inclusive (start, stop) of block b in 1..m (src/partitioning.jl:332-336)."""
    return load_balanced_endpoint(b - 1, n, m) + 1, load_balanced_endpoint(b, n, m)
