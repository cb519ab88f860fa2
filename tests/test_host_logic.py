"""Setup-time host logic of the library that needs no GPU: partitioning and cell ordering (C ABI, host code paths only)."""
import numpy as np


def test_process_partition_reference_goldens(J, O):
    # test/partitioning.jl:57-70
    N = O.cart_neighbors(5, 1, 1)
    assert J.process_partition(N, 5, [1, 1, 2, 1, 1]).tolist() == [1, 1, 2, 3, 3]
    p = np.array([1, 1, 2, 3, 3])
    out = J.process_partition(N, 5, p)
    assert out.tolist() == p.tolist() and out is not p
    assert J.process_partition(N, 5, p, weights=[1.0, 1.0, 1.0, 0.0]).tolist() == [1, 1, 2, 3, 4]
    # agrees with the oracle's restatement on a random partition of a 3-D grid
    w = J.workloads.unstructured_hex(6, 5, 4)
    part = np.random.default_rng(0).integers(1, 5, w["nc"])
    assert np.array_equal(J.process_partition(w["N"], w["nc"], part), O.process_partition(w["N"], w["nc"], part))


def test_partition_features(J):
    # test/partitioning.jl:11-36: every block non-empty, labels 1..np
    w = J.workloads.unstructured_hex(8, 7, 6)
    for k in range(1, 9):
        for kind in ("metis", "linear"):
            p = J.partition(w["N"], k, nc=w["nc"], partitioner=kind, weights=w["Tf"] if kind == "metis" else None)
            assert p.min() == 1 and p.max() == k and len(np.unique(p)) == k
    l = np.arange(1, 50); r = np.arange(2, 51)
    p = J.partition(np.stack([l, r], axis=1), 5)
    assert p.min() == 1 and p.max() == 5 and len(np.unique(p)) == 5
    # METIS blocks of a connected grid stay (nearly) connected: process_partition adds few blocks
    p8 = J.partition(w["N"], 8, nc=w["nc"])
    assert J.process_partition(w["N"], w["nc"], p8).max() <= 14


def test_multicolor_ordering_is_a_proper_colouring(J):
    w = J.workloads.unstructured_hex(7, 6, 5)
    perm, ncol = J.multicolor_ordering(w["N"], w["nc"])
    assert sorted(perm.tolist()) == list(range(1, w["nc"] + 1)) and ncol == 2
    N2 = perm[w["N"] - 1]
    n_first = np.sum(perm <= (w["nc"] + 1) // 2)
    # colour = label range; no face inside one range (proper colouring) for the two colour classes found
    sizes = [int(np.sum((N2[:, 0] <= k) & (N2[:, 1] <= k))) for k in range(1, w["nc"] + 1)]
    split = max(k for k in range(1, w["nc"] + 1) if sizes[k - 1] == 0)
    assert split >= w["nc"] // 2 - 1
    assert np.all((N2[:, 0] <= split) != (N2[:, 1] <= split))
