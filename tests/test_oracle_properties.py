"""Property tests of the checker's host logic on generated inputs (hypothesis, derandomised so that every run sees the same
examples): the structural invariants the reference's data structures guarantee, over ragged / duplicated / disconnected
inputs the fixed meshes of the other tests do not reach — and, in the last section, the library's host-only entry points
(partitioning, cell ordering: no device is touched) against the checker on the same generated inputs. No GPU."""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.csgraph as csg
from hypothesis import given, settings, strategies as st

SET = settings(derandomize=True, max_examples=60, deadline=None)


@st.composite
def graphs(draw, max_cells=14, max_faces=30):
    nc = draw(st.integers(2, max_cells))
    nf = draw(st.integers(0, max_faces))
    pairs = draw(st.lists(st.tuples(st.integers(1, nc), st.integers(1, nc)).filter(lambda t: t[0] != t[1]), min_size=nf, max_size=nf))
    return nc, np.array(pairs, dtype=np.int64).reshape(-1, 2)


@SET
@given(st.integers(1, 12), st.lists(st.tuples(st.integers(1, 12), st.integers(1, 12)), max_size=60))
def test_csr_from_coo_is_canonical(O, n, entries):
    """sparse(I, J, V) semantics (src/StaticCSR/mat.jl:73-76): rows complete, columns strictly ascending, duplicates merged."""
    entries = [(i, j) for i, j in entries if i <= n and j <= n]
    I = np.array([e[0] for e in entries], dtype=np.int64); Jc = np.array([e[1] for e in entries], dtype=np.int64)
    rowptr, colidx = O.csr_from_coo(I, Jc, n)
    assert rowptr[0] == 1 and rowptr[-1] == colidx.shape[0] + 1 and np.all(np.diff(rowptr) >= 0)
    got = set()
    for r in range(n):
        cols = colidx[rowptr[r] - 1:rowptr[r + 1] - 1]
        assert np.all(np.diff(cols) > 0)
        got |= {(r + 1, int(c)) for c in cols}
    assert got == set(entries)


@SET
@given(graphs())
def test_half_face_map_invariants(O, g):
    """get_cell_faces / half_face_map (src/utils.jl:813-841, src/domains.jl:101-122) on multigraphs with isolated cells: every
    face appears once from each side with signs +1 (left) / -1 (right), per-cell lists are sorted by face, and
    half_face_map_to_neighbors (src/domains.jl:124-137) recovers N."""
    nc, N = g
    hf = O.half_face_map(N, nc)
    pos, faces, cells, sign = hf["face_pos"], hf["faces"], hf["cells"], hf["face_sign"]
    assert pos[0] == 1 and pos[-1] == 2 * N.shape[0] + 1
    seen = {}
    for c in range(1, nc + 1):
        fl = faces[pos[c - 1] - 1:pos[c] - 1]
        assert np.all(np.diff(fl) >= 0)
        for k in range(pos[c - 1] - 1, pos[c] - 1):
            f = faces[k]
            left, right = N[f - 1]
            assert (sign[k] == 1 and c == left and cells[k] == right) or (sign[k] == -1 and c == right and cells[k] == left)
            seen.setdefault(int(f), []).append(int(sign[k]))
    assert all(sorted(v) == [-1, 1] for v in seen.values()) and len(seen) == N.shape[0]


@SET
@given(graphs(), st.integers(1, 4), st.integers(0, 2**31 - 1))
def test_process_partition_blocks_are_connected(O, g, nblocks, seed):
    """process_partition (src/partitioning.jl:139-180): every output block is connected in the face graph, no output block
    spans two input blocks, and input blocks that were connected keep their label."""
    nc, N = g
    part = np.random.default_rng(seed).integers(1, nblocks + 1, nc).astype(np.int64)
    out = O.process_partition(N, nc, part)
    A = sp.coo_matrix((np.ones(N.shape[0]), (N[:, 0] - 1, N[:, 1] - 1)), shape=(nc, nc)) if N.shape[0] else sp.coo_matrix((nc, nc))
    A = (A + A.T).tocsr()
    for b in np.unique(out):
        idx = np.flatnonzero(out == b)
        assert np.unique(part[idx]).shape[0] == 1
        ncomp, _ = csg.connected_components(A[idx][:, idx], directed=False)
        assert ncomp == 1
    for b in np.unique(part):
        idx = np.flatnonzero(part == b)
        ncomp, _ = csg.connected_components(A[idx][:, idx], directed=False)
        if ncomp == 1:
            assert np.all(out[idx] == b)
    assert np.unique(out).shape[0] >= np.unique(part).shape[0]


@SET
@given(st.integers(1, 9), st.integers(1, 200))
def test_partition_linear_is_balanced_and_monotone(O, m, n):
    """partition_linear (src/partitioning.jl:103-114): labels 1..m ascending, block sizes differ by at most one."""
    if m > n:
        m = n
    p = O.partition_linear(m, n)
    assert p[0] == 1 and p[-1] == m and np.all(np.diff(p) >= 0) and np.all(np.diff(p) <= 1)
    sizes = np.bincount(p)[1:]
    assert sizes.max() - sizes.min() <= 1


fl = st.floats(-1e3, 1e3, allow_nan=False, allow_infinity=False)


@SET
@given(st.lists(st.tuples(st.floats(0.0, 1.0), fl), min_size=1, max_size=20), st.floats(0.01, 1.0), st.floats(0.1, 1.0))
def test_fraction_pair_update_stays_feasible(O, cells, abs_max, w):
    """unit_update_pairs! (src/variables/utils.jl:471-482): after the update both fractions lie in [0, 1], they sum to one, and
    the first moved by at most abs_max in the direction of its increment."""
    s = np.array([[a, 1.0 - a] for a, _ in cells]).ravel()
    dx = np.array([d for _, d in cells])
    s0 = s.copy()
    O.update_fraction_pair(s, dx, w=w, abs_max=abs_max)
    a0, a1, b1 = s0[0::2], s[0::2], s[1::2]
    assert np.all(a1 >= 0) and np.all(a1 <= 1) and np.all(b1 >= 0) and np.all(b1 <= 1)
    assert np.allclose(a1 + b1, 1.0, rtol=0, atol=4e-16)
    assert np.all(np.abs(a1 - a0) <= abs_max * (1 + 1e-12))
    assert np.all((a1 - a0) * dx >= 0)


@SET
@given(st.lists(st.tuples(st.floats(1.0, 100.0), fl), min_size=1, max_size=20), st.floats(0.1, 10.0), st.floats(0.01, 0.5))
def test_scalar_update_respects_every_limit(O, cells, abs_max, rel_max):
    """update_jutul_variable_internal! / choose_increment (src/variables/utils.jl:117-174): absolute and relative increment
    limits and the value bounds all hold at once; an increment inside every limit is applied unchanged."""
    v = np.array([a for a, _ in cells]); dx = np.array([d for _, d in cells])
    v0 = v.copy()
    O.update_scalar(v, dx, abs_max=abs_max, rel_max=rel_max, minv=0.5, maxv=150.0)
    dv = v - v0
    assert np.all(np.abs(dv) <= abs_max * (1 + 1e-12)) and np.all(np.abs(dv) <= rel_max * np.abs(v0) * (1 + 1e-12) + 1e-300)
    assert np.all(v >= 0.5) and np.all(v <= 150.0) and np.all(dv * dx >= 0)
    free = (np.abs(dx) <= abs_max) & (np.abs(dx) <= rel_max * np.abs(v0)) & (v0 + dx >= 0.5) & (v0 + dx <= 150.0)
    assert np.array_equal(v[free], (v0 + dx)[free])


# ---- the library's own host-only entry points (no GPU is touched: partition.cu host code through the C ABI) against the checker
@SET
@given(graphs(), st.integers(1, 4), st.integers(0, 2**31 - 1), st.booleans())
def test_library_process_partition_equals_checker(J, O, g, nblocks, seed, weighted):
    """jb_process_partition == the restatement of process_partition, label for label, on multigraphs with isolated cells and
    zero-weight faces (weights == 0 cut the connection, src/partitioning.jl:139-160)."""
    nc, N = g
    rng = np.random.default_rng(seed)
    part = rng.integers(1, nblocks + 1, nc).astype(np.int64)
    wts = (rng.random(N.shape[0]) > 0.3).astype(np.float64) if weighted else None
    assert np.array_equal(J.process_partition(N, nc, part, weights=wts), O.process_partition(N, nc, part, weights=wts))


@SET
@given(graphs(max_cells=40, max_faces=120))
def test_library_multicolor_ordering_is_always_a_proper_colouring(J, g):
    """jb_order_multicolor on arbitrary multigraphs: a permutation of 1..nc whose labels split into `ncolors` consecutive ranges
    with no face inside a range — the property the two-colour ILU(0) sweeps and the identity rows rely on."""
    nc, N = g
    perm, ncol = J.multicolor_ordering(N, nc)
    assert sorted(perm.tolist()) == list(range(1, nc + 1)) and 1 <= ncol <= nc
    if N.shape[0] == 0:
        return
    a, b = perm[N[:, 0] - 1], perm[N[:, 1] - 1]
    # greedy reconstruction of the colour ranges: a new range starts at the first label adjacent to a label of the current range
    lo = np.minimum(a, b); hi = np.maximum(a, b)
    starts, cur = [1], 1
    for lab in range(2, nc + 1):
        if np.any((hi == lab) & (lo >= cur)):
            starts.append(lab); cur = lab
    assert len(starts) <= ncol


@SET
@given(st.integers(2, 7), st.integers(2, 6), st.integers(1, 5), st.integers(1, 6))
def test_library_partitions_cover_all_labels(J, nx, ny, nz, k):
    """partition(N, k) (src/partitioning.jl:244-307; test/partitioning.jl:11-36): labels 1..k, every block non-empty, for METIS
    and linear partitioners on small grids, k up to the number of cells. METIS may leave a block empty when k approaches the
    number of cells; the library then fails loudly ("a block is empty") instead of returning a partition with a hole."""
    w = J.workloads.unstructured_hex(nx, ny, nz)
    k = min(k, w["nc"])
    for kind in ("metis", "linear"):
        try:
            p = J.partition(w["N"], k, nc=w["nc"], partitioner=kind)
        except J.JutulB200Error as e:
            assert kind == "metis" and 4 * k > w["nc"] and "a block is empty" in str(e)
            continue
        assert p.shape[0] == w["nc"] and set(np.unique(p).tolist()) == set(range(1, k + 1))
