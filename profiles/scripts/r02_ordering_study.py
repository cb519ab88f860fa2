"""Round 2: how does the cell numbering change the ILU(0)-BiCGStab iteration count?  (CPU, oracle only.)

The two-phase problem of SURVEY §8(d) on an UN-permuted nx x ny x nz hex mesh (cell = i + nx (j + ny k)), first Newton
iteration of report step 1, un-partitioned ILU(0), right-preconditioned BiCGStab. Orderings:
  natural     ijk numbering as the mesh generator emits it (what the reference sees for a structured input)
  random      the benchmark's seeded random relabelling (what the reference sees for bench.py's mesh)
  rcm         breadth-first Cuthill-McKee from cell 1 (reversed)
  colour k    colour = (i + j + k) mod k, numbered colour by colour (k = 2 is the red-black ordering of the device path)
  multicolor  jb_order_multicolor (the device path's own renumbering: BFS locality + greedy colouring)
Output: one line per (ordering, rtol): ILU levels, linear iterations, final relative residual.
usage: python profiles/scripts/r02_ordering_study.py [n] [rtol ...]
"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import __graft_entry__ as g  # noqa: E402

J = g.load_package()
import oracle as O  # noqa: E402
from conftest import oracle_system  # noqa: E402


def relabel(w, perm):
    """perm[c] = new 1-based label of 1-based cell c+1."""
    n = w["nc"]
    wp = dict(w)
    wp["N"] = perm[w["N"] - 1]
    inv = np.empty(n, dtype=np.int64); inv[perm - 1] = np.arange(n)
    for k in ("pv", "p0", "sw0"):
        wp[k] = w[k][inv]
    wp["src_cells"] = perm[np.asarray(w["src_cells"], dtype=np.int64) - 1]
    return wp


def rcm(N, n):
    import scipy.sparse as sp
    from scipy.sparse.csgraph import reverse_cuthill_mckee
    A = sp.coo_matrix((np.ones(N.shape[0]), (N[:, 0] - 1, N[:, 1] - 1)), shape=(n, n)).tocsr()
    A = A + A.T
    order = reverse_cuthill_mckee(A, symmetric_mode=True)
    perm = np.empty(n, dtype=np.int64); perm[order] = np.arange(1, n + 1)
    return perm


def colour_perm(nx, ny, nz, k):
    n = nx * ny * nz
    kk, jj, ii = np.unravel_index(np.arange(n), (nz, ny, nx))
    col = (ii + jj + kk) % k
    order = np.lexsort((np.arange(n), col))
    perm = np.empty(n, dtype=np.int64); perm[order] = np.arange(1, n + 1)
    return perm


def run(w, rtols, label):
    n = w["nc"]
    s = oracle_system(O, w)
    M0 = O.mass_2ph(w["pv"], w["params"], w["p0"], w["sw0"])
    nz, r = O.assemble_2ph(s["hf"], s["diag_pos"], s["hf_pos"], w["Tf"], w["gdz"], w["pv"], w["params"], w["p0"], w["sw0"], M0, w["dt"],
                           s["colidx"].shape[0], w["src_cells"], w["src_vals"])
    ilu = O.ILU0(n, 2, s["rowptr"], s["colidx"])
    ilu.factor(nz)
    nlev = ilu.set_level_schedule(True, max_levels=10 ** 9)
    if nlev > 64:
        ilu.set_level_schedule(False)
    for rtol in rtols:
        t0 = time.time()
        x, st, its, hist = O.bicgstab(n, 2, s["rowptr"], s["colidx"], nz, r, ilu, rtol=rtol, itmax=2000)
        print(f"{label:12s} rtol {rtol:7.0e}  levels {nlev:5d}  iterations {its:5d}  status {st}  rel.res {hist[its] / hist[0]:.2e}  ({time.time() - t0:.1f} s)",
              flush=True)


def main():
    nn = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    rtols = [float(a) for a in sys.argv[2:]] or [1e-3, 1e-6]
    w = J.workloads.unstructured_hex(nn, nn, nn, permute=False)
    n = w["nc"]
    print(f"# {nn}^3 = {n} cells, un-permuted mesh; first Newton iteration of report step 1")
    run(w, rtols, "natural")
    run(relabel(w, np.random.default_rng(20261017).permutation(n) + 1), rtols, "random")
    run(relabel(w, rcm(w["N"], n)), rtols, "rcm")
    for k in (2, 3, 4, 6, 8, 16):
        run(relabel(w, colour_perm(nn, nn, nn, k)), rtols, f"colour {k}")
    perm, ncol = J.multicolor_ordering(w["N"], n)
    run(relabel(w, perm), rtols, f"multicolor{ncol}")


if __name__ == "__main__":
    main()
