"""Iteration counts of ILU(0)-BiCGStab under different cell numberings (CPU oracle): VERDICT item 6."""
import sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import __graft_entry__ as g
J = g.load_package()
import oracle as O

def renumber(w, newlabel):
    w2 = dict(w)
    N = w["N"] - 1
    w2["N"] = newlabel[N] + 1
    for k in ("pv", "p0", "sw0", "z"):
        a = np.empty_like(w[k]); a[newlabel] = w[k]; w2[k] = a
    w2["src_cells"] = newlabel[w["src_cells"] - 1] + 1
    return w2

def system(w):
    hf = O.half_face_map(w["N"], w["nc"])
    I, Jc = O.tpfa_pattern(hf)
    rowptr, colidx = O.csr_from_coo(I, Jc, w["nc"])
    dpos, hpos = O.tpfa_alignment(hf, rowptr, colidx)
    M0 = O.mass_2ph(w["pv"], w["params"], w["p0"], w["sw0"])
    nz, r = O.assemble_2ph(hf, dpos, hpos, w["Tf"], w["gdz"], w["pv"], w["params"], w["p0"], w["sw0"], M0, w["dt"], colidx.shape[0], w["src_cells"], w["src_vals"])
    return rowptr, colidx, nz, r

def levels(rowptr, colidx, n):
    """forward dependency levels of the natural-order sweep"""
    lev = np.zeros(n, dtype=np.int64)
    rp = rowptr - 1; ci = colidx - 1
    for i in range(n):
        cols = ci[rp[i]:rp[i + 1]]
        low = cols[cols < i]
        lev[i] = (lev[low].max() + 1) if low.size else 0
    return int(lev.max()) + 1

dims = tuple(int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (48, 48, 48)
wn = J.workloads.unstructured_hex(*dims, permute=False)
nc = wn["nc"]
rng = np.random.default_rng(5)
k, j, i = np.unravel_index(np.arange(nc), (dims[2], dims[1], dims[0]))
def by_key(key):
    return np.argsort(np.argsort(key, kind="stable"), kind="stable")
orders = {"natural": np.arange(nc), "random": rng.permutation(nc), "redblack": by_key(((i + j + k) % 2) * nc + np.arange(nc))}
orders["color8"] = by_key(((i % 2) + 2 * (j % 2) + 4 * (k % 2)) * nc + np.arange(nc))
for b in (2, 4, 8):
    bi, bj, bk = i // b, j // b, k // b
    nbx, nby = (dims[0] + b - 1) // b, (dims[1] + b - 1) // b
    blk = bi + nbx * (bj + nby * bk)
    col = (bi + bj + bk) % 2
    inside = (i % b) + b * ((j % b) + b * (k % b))
    orders[f"blockRB{b}"] = by_key((col * (blk.max() + 1) + blk) * (b ** 3) + inside)
import scipy.sparse as sp
from scipy.sparse.csgraph import reverse_cuthill_mckee
N0 = wn["N"] - 1
G = sp.coo_matrix((np.ones(N0.shape[0]), (N0[:, 0], N0[:, 1])), shape=(nc, nc)); G = (G + G.T).tocsr()
rcm = reverse_cuthill_mckee(G, symmetric_mode=True)
lab = np.empty(nc, dtype=np.int64); lab[rcm] = np.arange(nc)
orders["rcm"] = lab
print("dims", dims, "threads", O.num_threads(), flush=True)
for name, lab in orders.items():
    w2 = renumber(wn, lab)
    rowptr, colidx, nz, r = system(w2)
    nlev = levels(rowptr, colidx, nc) if nc <= 300000 else -1
    ilu = O.ILU0(nc, 2, rowptr, colidx, None); ilu.factor(nz)
    out = []
    for rtol in (1e-3, 1e-6):
        x, st, its, hist = O.bicgstab(nc, 2, rowptr, colidx, nz, r, ilu, rtol=rtol, itmax=1000)
        out.append((rtol, its, st))
    print(f"{name:10s} levels {nlev:5d}  " + "  ".join(f"rtol {a:g}: {b} its (st {c})" for a, b, c in out), flush=True)
