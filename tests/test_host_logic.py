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


def test_decompose_numbers_boundary_cells_last_in_each_colour(J, monkeypatch):
    """Distributed local numbering (dist.decompose, multicolour order): owned cells first and coloured properly; owned
    cells that touch a ghost sit at the END of their colour, so the rows before them form contiguous runs without ghost
    couplings — what lets whole stream chunks qualify as identity rows of the right-preconditioned operator on every rank
    (csrc/krylov.cu) and what the interior/boundary SpMV split uses."""
    from jutul_b200 import dist as D
    monkeypatch.delenv("JB_RB_IDENTITY", raising=False)
    monkeypatch.delenv("JB_OVERLAP", raising=False)
    w = J.workloads.unstructured_hex(12, 10, 8)
    nc = w["nc"]
    part = J.partition(w["N"], 3, weights=w["Tf"], nc=nc)
    for rank in range(3):
        plan = D.decompose(w["N"], nc, part, rank, "multicolor")
        no, nl = plan["n_owned"], plan["n_local"]
        assert np.array_equal(np.sort(plan["owned"]), np.nonzero(np.asarray(part) - 1 == rank)[0])
        Nl = plan["N_local"] - 1
        owned_face = (Nl[:, 0] < no) & (Nl[:, 1] < no)
        touches_ghost = np.zeros(nl, dtype=bool)
        gf = ~owned_face
        touches_ghost[Nl[gf, 0]] = True; touches_ghost[Nl[gf, 1]] = True
        touches_ghost[no:] = False
        # colours: maximal label ranges without an internal owned face
        lo = np.minimum(Nl[owned_face, 0], Nl[owned_face, 1]); hi = np.maximum(Nl[owned_face, 0], Nl[owned_face, 1])
        ncol = plan["ncolors"]
        assert ncol is not None and ncol >= 2
        # inside every colour the boundary cells come last: once a ghost-touching owned cell appears, no interior cell follows
        # (checked on the colour ranges reported by the proper-colouring property: no owned face inside a range)
        bounds = sorted(set([0, no] + [int(k) for k in np.nonzero(np.diff(touches_ghost[:no].astype(int)) == -1)[0] + 1]))
        assert len(bounds) - 1 <= ncol + 1        # at most one interior->boundary->interior transition per colour
        for a, b in zip(bounds[:-1], bounds[1:]):
            seg = touches_ghost[a:b]
            first_b = int(np.argmax(seg)) if seg.any() else len(seg)
            assert np.all(seg[first_b:])          # boundary cells form the tail of the segment
            inside = (lo >= a) & (hi < b)
            assert not inside.any()               # and the segment is a colour: no owned face inside it
    monkeypatch.setenv("JB_RB_IDENTITY", "0")
    plan0 = D.decompose(w["N"], nc, part, 0, "multicolor")
    assert np.array_equal(np.sort(plan0["owned"]), np.sort(D.decompose(w["N"], nc, part, 0, "default")["owned"]))


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (CPU restatement on the host cores) prints ONE JSON line that carries the keys of the
    bench contract, on the same metric / unit / config keys as the GPU arm; ranks != 0 print nothing."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--dims", "12,10,8", "--steps", "1", "--warmup", "0",
           "--reference-budget", "5"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "newton_iterations_per_sec" and d["unit"] == "newton_it/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f64" and "workload" in d["config"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["value"] > 0 and d["newton_iterations_per_step"] >= 1
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out1 = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=root, env=env)
    assert out1.returncode == 0 and out1.stdout.strip() == ""


def test_interpolator_capping_matches_oracle(J):
    """get_1d_interpolator / get_2d_interpolator end-point capping of the host mirror (jutul.jl_b200/variables.py) equals the
    oracle restatement of src/interpolation.jl:118-154,224-285 (the device only ever sees the capped arrays)."""
    from oracle import widen as W
    V = J.variables
    xs = np.array([0.0, 0.5, 1.5, 4.0]); ys = xs ** 2
    X, F = V.cap_1d(xs, ys)
    I = W.get_1d_interpolator(xs, ys)
    assert np.array_equal(X, I.X) and np.array_equal(F, I.F)
    X1, F1 = V.cap_1d([3.0], [7.0])
    I1 = W.get_1d_interpolator([3.0], [7.0])
    assert np.array_equal(X1, I1.X) and np.array_equal(F1, I1.F)
    gx = np.array([0.0, 1.0, 3.0]); gy = np.array([-1.0, 0.0, 2.0, 5.0])
    fs = np.arange(12.0).reshape(3, 4)
    X2, Y2, F2 = V.cap_2d(gx, gy, fs)
    I2 = W.get_2d_interpolator(gx, gy, fs)
    assert np.array_equal(X2, I2.X) and np.array_equal(Y2, I2.Y) and np.array_equal(F2, I2.F)
    # VarSpec mirrors jb_var_spec: 4 + 12 + 4 + 4 + 32 bytes, doubles 8-byte aligned
    import ctypes as C
    assert C.sizeof(V.VarSpec) == 56 and V.VarSpec.c.offset == 24
