"""Synthetic inputs for the Newton hot path. numpy only; no oracle code.

`unstructured_hex(nx, ny, nz)` builds a hexahedral Cartesian topology in the reference's MRST
face order (src/meshes/cart.jl:197-225), then makes it genuinely unstructured: a seeded
permutation of the cell numbers, a seeded permutation of the face numbers and random left/right
swaps, so no kernel can exploit ijk structure. Two-phase rock/fluid data follow §8(d).

Random streams: numpy `default_rng(seed)` (PCG64), seed 20261017 unless given; draws are made
in a fixed order (cell permutation, face permutation, flips, permeability, porosity, Sw phases).
"""
import numpy as np

GRAVITY = 9.80665
MILLIDARCY = 9.869233e-16
DAY = 86400.0
DEFAULT_SEED = 20261017


def cart_neighbors(nx, ny=1, nz=1):
    """0-based (nf, 2) neighbour list in MRST order."""
    idx = np.arange(nx * ny * nz, dtype=np.int64).reshape(nz, ny, nx)
    xl = idx[:, :, :-1].ravel()
    yl = np.transpose(idx[:, :-1, :], (1, 0, 2)).ravel()
    zl = idx[:-1, :, :].ravel()
    left = np.concatenate([xl, yl, zl])
    right = np.concatenate([xl + 1, yl + nx, zl + nx * ny])
    direction = np.concatenate([np.zeros(xl.size, np.int8), np.ones(yl.size, np.int8), np.full(zl.size, 2, np.int8)])
    return np.stack([left, right], axis=1), direction


def unstructured_hex(nx, ny, nz, cell_size=(10.0, 10.0, 2.0), seed=DEFAULT_SEED, permute=True, jitter_perm=True):
    """Returns a dict with N (nf x 2, 1-based int64 = Julia's 2 x nf), nc, nf and two-phase data:
    Tf, gdz (per face), pv (per cell), p0/sw0 initial state, sources, params, dt."""
    rng = np.random.default_rng(seed)
    nc = nx * ny * nz
    N0, direction = cart_neighbors(nx, ny, nz)
    nf = N0.shape[0]
    dx, dy, dz = cell_size
    d = np.array([dx, dy, dz])
    area = np.array([dy * dz, dx * dz, dx * dy])
    vol = dx * dy * dz
    # cell data in ijk numbering
    k_idx, j_idx, i_idx = np.unravel_index(np.arange(nc), (nz, ny, nx))
    z = (k_idx + 0.5) * dz
    if permute:
        cell_perm = rng.permutation(nc)          # new label of old cell
        face_perm = rng.permutation(nf)          # new label of old face
        flip = rng.random(nf) < 0.5
    else:
        cell_perm = np.arange(nc); face_perm = np.arange(nf); flip = np.zeros(nf, bool)
    perm_K = np.exp(rng.uniform(np.log(10.0), np.log(1000.0), nc)) * MILLIDARCY if jitter_perm else np.full(nc, 100.0 * MILLIDARCY)
    poro = rng.uniform(0.1, 0.3, nc)
    ph = rng.uniform(0, 2 * np.pi, 3)
    sw = 0.5 + 0.3 * np.sin(2 * np.pi * i_idx / max(nx, 2) * 1.5 + ph[0]) * np.cos(2 * np.pi * j_idx / max(ny, 2) + ph[1]) \
        * np.cos(np.pi * k_idx / max(nz, 2) + ph[2])
    p_init = 1e7 + 1000.0 * GRAVITY * z
    # TPFA transmissibility: half_face_trans = A*K/(d/2) on an orthogonal grid (finite-volume.jl:220-222),
    # harmonic combination (compute_face_trans :224-233); gdz = -g (z_r - z_l) (:304-313)
    l0, r0 = N0[:, 0], N0[:, 1]
    A = area[direction]; half = 0.5 * d[direction]
    Tl = A * perm_K[l0] / half; Tr = A * perm_K[r0] / half
    Tf0 = 1.0 / (1.0 / Tl + 1.0 / Tr)
    # relabel
    Lnew = cell_perm[l0]; Rnew = cell_perm[r0]
    Lf = np.where(flip, Rnew, Lnew); Rf = np.where(flip, Lnew, Rnew)
    N = np.empty((nf, 2), dtype=np.int64)
    N[face_perm, 0] = Lf + 1
    N[face_perm, 1] = Rf + 1
    Tf = np.empty(nf); Tf[face_perm] = Tf0

    def cells_to_new(a):
        out = np.empty_like(a); out[cell_perm] = a
        return out

    z_new = cells_to_new(z)
    gdz = -GRAVITY * (z_new[N[:, 1] - 1] - z_new[N[:, 0] - 1])
    pv = cells_to_new(poro * vol)
    inj = int(cell_perm[0]) + 1
    prod = int(cell_perm[nc - 1]) + 1
    # kg/s. The producer takes a fixed mass rate of EACH phase (a constant source on the diagonal entries, as the reference's
    # apply_forces! hook models it): 0.0025 kg/s of oil is ~1.5 % of the cell's oil per day, so with the inflow from its
    # neighbours the cell keeps oil through the 50 timesteps of config 5 (with 0.02 it ran dry after 26 steps and the Newton
    # loop could no longer satisfy the oil sink)
    q = 0.005
    return dict(
        nx=nx, ny=ny, nz=nz, nc=nc, nf=nf, N=N, Tf=Tf, gdz=gdz, pv=pv, z=z_new,
        p0=cells_to_new(p_init), sw0=cells_to_new(sw), cell_perm=cell_perm, face_perm=face_perm,
        params=np.array([1000.0, 700.0, 4.5e-10, 1e-9, 1e-3, 5e-3, 1e5]),
        src_cells=np.array([inj, prod], dtype=np.int64), src_vals=np.array([[-q, 0.0], [0.5 * q, 0.5 * q]]),
        dt=DAY, seed=seed,
    )


def algorithmic_bytes(nc, nf, bs=2):
    """Algorithmic HBM bytes per launch of each kernel (BASELINE.md §3 / SURVEY.md §8(d))."""
    nb = nc + 2 * nf
    spmv = nb * (8 * bs * bs + 4) + 4 * (nc + 1) + 2 * nc * bs * 8
    assembly = nb * 8 * bs * bs + nc * bs * 8 + nc * (2 * bs + 1) * 8 + nf * 24 + nb * 4
    ilu_apply = (nb - nc) * (8 * bs * bs + 4) + nc * 8 * bs * bs + 8 * (nc + 1) + 2 * nc * bs * 8
    ilu_factor = 2 * nb * 8 * bs * bs
    bicgstab_iter = 2 * spmv + 2 * ilu_apply + 19 * nc * bs * 8
    return dict(spmv=spmv, assembly=assembly, ilu_apply=ilu_apply, ilu_factor=ilu_factor, bicgstab_iter=bicgstab_iter)


def fused_iteration_bytes(nb, n_local, n_own, n_id, id_blocks, n_b, n_l, n_u):
    """Compulsory HBM bytes of ONE launch of the fused BiCGStab iteration kernel (csrc/krylov_persistent.cu), phase by phase:
    every matrix block (32 B + 4 B column) and every vector element (8 B) counted once per phase that must stream it;
    gathered operands counted once per phase (re-reads through L2 are the cache-efficiency loss the fraction exposes).
    nb blocks of the local Jacobian, n_own rows updated, n_id identity rows holding id_blocks blocks, n_b rows of the second
    colour, n_l / n_u blocks of L / U. bs = 2 (16 B per row and vector)."""
    n_r = n_own - n_b                       # rows swept with U (first colour)
    nb_a = nb - id_blocks - (n_local - n_own)   # blocks of the rows the SpMV phases stream (ghost rows are never streamed)
    nr_a = n_own - n_id
    ph = {}
    ph["A1_sweep_L_y"] = n_l * 36 + n_b * (8 + 32 + 80) + 16 * n_r
    ph["A2_sweep_U_y"] = n_u * 36 + n_r * 72 + 16 * n_id + 16 * n_b
    ph["A3_spmv_v"] = nb_a * 36 + nr_a * 36 + 16 * n_local
    ph["V1_s"] = 48 * n_r
    ph["A4_sweep_L_z"] = n_l * 36 + n_b * 104 + 16 * n_r
    ph["A5_sweep_U_z"] = n_u * 36 + n_r * 72 + 16 * n_b
    ph["A6_spmv_t"] = nb_a * 36 + nr_a * 36 + 16 * n_local
    ph["V2_x_r"] = 112 * n_own + 16 * nr_a
    ph["V3_p"] = 48 * n_r + 16 * max(n_r - n_id, 0)
    ph["total"] = sum(ph.values())
    return ph


def heat_initial_condition(nx, ny):
    """docs/src/index.md:22-46: T0 = 100 inside the centred half-width square, else 0."""
    T = np.zeros((ny, nx))
    i = np.arange(nx); j = np.arange(ny)
    x = (i + 0.5) / nx; y = (j + 0.5) / ny
    inside = (np.abs(x - 0.5)[None, :] < 0.25) & (np.abs(y - 0.5)[:, None] < 0.25)
    T[inside] = 100.0
    return T.ravel()
