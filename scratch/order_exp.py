import sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import __graft_entry__ as g
J = g.load_package()
import oracle as O

def renumber(w, newlabel):
    """newlabel[c_old0] = new 0-based label"""
    w2 = dict(w)
    N = w["N"] - 1
    w2["N"] = newlabel[N] + 1
    for k in ("pv", "p0", "sw0", "z"):
        a = np.empty_like(w[k]); a[newlabel] = w[k]; w2[k] = a
    w2["src_cells"] = newlabel[w["src_cells"] - 1] + 1
    return w2

def solve_its(w, rtol, part=None):
    hf = O.half_face_map(w["N"], w["nc"])
    I, Jc = O.tpfa_pattern(hf)
    rowptr, colidx = O.csr_from_coo(I, Jc, w["nc"])
    dpos, hpos = O.tpfa_alignment(hf, rowptr, colidx)
    M0 = O.mass_2ph(w["pv"], w["params"], w["p0"], w["sw0"])
    nz, r = O.assemble_2ph(hf, dpos, hpos, w["Tf"], w["gdz"], w["pv"], w["params"], w["p0"], w["sw0"], M0, w["dt"], colidx.shape[0], w["src_cells"], w["src_vals"])
    ilu = O.ILU0(w["nc"], 2, rowptr, colidx, part); ilu.factor(nz)
    x, st, its, hist = O.bicgstab(w["nc"], 2, rowptr, colidx, nz, r, ilu, rtol=rtol, itmax=1000)
    return its, st

dims = (40, 40, 40)
wn = J.workloads.unstructured_hex(*dims, permute=False)   # natural ijk numbering
nc = wn["nc"]
rng = np.random.default_rng(5)
k, j, i = np.unravel_index(np.arange(nc), (dims[2], dims[1], dims[0]))
orders = {
    "natural": np.arange(nc),
    "random": rng.permutation(nc),
    "redblack": np.argsort(np.argsort(((i + j + k) % 2) * nc + np.arange(nc), kind="stable")),
}
# 8-colour (2x2x2 parity) and "diagonal wavefront blocks"
col8 = (i % 2) + 2 * (j % 2) + 4 * (k % 2)
orders["color8"] = np.argsort(np.argsort(col8 * nc + np.arange(nc), kind="stable"))
for name, lab in orders.items():
    w2 = renumber(wn, lab)
    for rtol in (1e-3, 1e-6):
        its, st = solve_its(w2, rtol)
        print(name, rtol, "its", its, "st", st, flush=True)
