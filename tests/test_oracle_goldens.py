"""Pins the CPU oracle against the reference's own golden vectors / known answers
(SURVEY.md §8(c)). No GPU."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

from conftest import to_scipy


def test_cart_neighbors_mrst_order(O, J):
    # src/meshes/cart.jl:197-225 by hand for 2x2: x-faces (1,2),(3,4); y-faces (1,3),(2,4)
    assert O.cart_neighbors(2, 2, 1).tolist() == [[1, 2], [3, 4], [1, 3], [2, 4]]
    # 3x2x2: first the 8 x-faces z,y,x order, then y-faces with loop order y,z,x, then z-faces
    N = O.cart_neighbors(3, 2, 2)
    assert N[:8].tolist() == [[1, 2], [2, 3], [4, 5], [5, 6], [7, 8], [8, 9], [10, 11], [11, 12]]
    assert N[8:14].tolist() == [[1, 4], [2, 5], [3, 6], [7, 10], [8, 11], [9, 12]]
    assert N[14:].tolist() == [[1, 7], [2, 8], [3, 9], [4, 10], [5, 11], [6, 12]]
    # the independent numpy restatement used by the workload generator agrees (test/mesh.jl:143-163 sizes)
    for dims in [(3, 1, 1), (3, 2, 1), (3, 2, 2), (4, 1, 1), (9, 7, 5), (100, 3, 7)]:
        N0, _ = J.workloads.cart_neighbors(*dims)
        assert np.array_equal(N0 + 1, O.cart_neighbors(*dims))


def test_half_face_map_known(O):
    # 3x1 mesh: faces (1,2),(2,3). get_cell_faces + half_face_map (src/utils.jl:813-841, src/domains.jl:101-122)
    hf = O.half_face_map(np.array([[1, 2], [2, 3]]), 3)
    assert hf["faces"].tolist() == [1, 1, 2, 2]
    assert hf["face_pos"].tolist() == [1, 2, 4, 5]
    assert hf["cells"].tolist() == [2, 1, 3, 2]
    assert hf["face_sign"].tolist() == [1, -1, 1, -1]
    # faces sorted ascending per cell even when N lists them out of order
    hf = O.half_face_map(np.array([[3, 2], [1, 2], [1, 3]]), 3)
    assert hf["faces"].tolist() == [2, 3, 1, 2, 1, 3]
    assert hf["cells"].tolist() == [2, 3, 3, 1, 2, 1]
    assert hf["face_sign"].tolist() == [1, 1, -1, -1, 1, -1]


def test_half_face_map_roundtrip(O):
    # half_face_map_to_neighbors (src/domains.jl:124-137) inverts the map
    N = O.cart_neighbors(4, 3, 2)
    hf = O.half_face_map(N, 24)
    N2 = np.zeros_like(N)
    for f, c, s in zip(hf["faces"], hf["cells"], hf["face_sign"]):
        N2[f - 1, 0 if s == -1 else 1] = c
    assert np.array_equal(N, N2)


def test_transmissibility_unit_cube(O):
    # test/utils.jl:293-298: on CartesianMesh((2,2,2)) with unit permeability every boundary half-trans is 1;
    # interior half-faces of the same cells use the same half_face_trans formula and geometry => also 1.
    geo = O.cart_geometry(2, 2, 2, 0.5, 0.5, 0.5)
    N = O.cart_neighbors(2, 2, 2)
    hf = O.half_face_map(N, 8)
    Thf = O.half_face_trans(geo, 1.0, hf)
    assert np.all(Thf == 1.0)
    T = O.face_trans(Thf, hf["faces"], N.shape[0])
    assert np.all(T == 0.5)


def test_workload_trans_matches_oracle(O, J):
    w = J.workloads.unstructured_hex(5, 4, 3, permute=False)
    geo = O.cart_geometry(5, 4, 3, 10.0, 10.0, 2.0)
    hf = O.half_face_map(w["N"], w["nc"])
    rng = np.random.default_rng(w["seed"])
    perm = np.exp(rng.uniform(np.log(10.0), np.log(1000.0), w["nc"])) * J.workloads.MILLIDARCY
    T = O.face_trans(O.half_face_trans(geo, perm, hf), hf["faces"], w["nf"])
    assert np.allclose(T, w["Tf"], rtol=1e-13)
    assert np.allclose(O.face_gdz(w["N"], geo["cell_centroids"][:, 2]), w["gdz"], rtol=1e-13, atol=1e-13)


def test_layout_goldens(O):
    # test/adjoints/utils.jl:57-68: equation-major vs block(entity)-major ordering of 7 dof on 2 cells
    import json, os
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_known_answers.json")))["layout_transfer"]
    eq_x_ref, block_y_ref = gold["eq_x_ref"], gold["block_y_ref"]
    nc, ndof = 2, 7
    for c in range(1, nc + 1):
        for d in range(1, ndof + 1):
            assert eq_x_ref[O.index_equation_major(c, d, nc) - 1] == block_y_ref[O.index_entity_major(c, d, ndof) - 1]
    # block nz index, src/equations.jl:111-114: column-major inside the block
    assert O.block_nz_index(1, 2, 1, 1) == 1 and O.block_nz_index(1, 2, 2, 1) == 2
    assert O.block_nz_index(1, 2, 1, 2) == 3 and O.block_nz_index(3, 2, 2, 2) == 12


def test_process_partition_goldens(O):
    # test/partitioning.jl:57-70
    N = O.cart_neighbors(5, 1, 1)
    assert O.process_partition(N, 5, [1, 1, 2, 1, 1]).tolist() == [1, 1, 2, 3, 3]
    assert O.process_partition(N, 5, [1, 1, 2, 3, 3]).tolist() == [1, 1, 2, 3, 3]
    assert O.process_partition(N, 5, [1, 1, 2, 3, 3], weights=[1.0, 1.0, 1.0, 0.0]).tolist() == [1, 1, 2, 3, 4]
    # partition_linear: blocks non-empty, min 1, max np (test/partitioning.jl:11-27)
    for npart in range(1, 11):
        p = O.partition_linear(npart, 100)
        assert p.min() == 1 and p.max() == npart and len(np.unique(p)) == npart


def test_poisson_known_answer(O):
    # test/test_systems/variable_poisson.jl:5-35: 3x1 unit square, coefficient 1, sources +1 / -1,
    # U - U[1] ≈ [0, 1/3, 2/3]. One Newton step with a direct solve (the reference default).
    nx, nc = 3, 3
    N = O.cart_neighbors(nx, 1, 1)
    hf = O.half_face_map(N, nc)
    geo = O.cart_geometry(nx, 1, 1, 1.0 / nx, 1.0, 1.0)
    K = O.face_trans(O.half_face_trans(geo, 1.0, hf), hf["faces"], N.shape[0])
    assert np.allclose(K, 3.0)
    I, Jc = O.tpfa_pattern(hf)
    rowptr, colidx = O.csr_from_coo(I, Jc, nc)
    U = np.ones(nc)
    for _ in range(2):
        nz, r = O.assemble_poisson(hf, rowptr, colidx, K, U, U, False, 1.0, [1, nc], [1.0, -1.0])
        if np.abs(r).max() <= 1e-3:
            break
        A = to_scipy(nc, 1, rowptr, colidx, nz)
        U = U - spla.spsolve(A.tocsc(), r)
    assert np.allclose(U - U[0], [0.0, 1.0 / 3.0, 2.0 / 3.0], atol=1e-6)


def test_heat_jacobian_and_residual_identity(O, J):
    # test/test_systems/helper.jl:3-18 (residual identity) + docs heat example pattern (heat_2d.jl:13-45)
    nx = ny = 6
    I, Jc = O.heat_pattern(nx, ny)
    rowptr, colidx = O.csr_from_coo(I, Jc, nx * ny)
    assert colidx.shape[0] == 5 * nx * ny
    rng = np.random.default_rng(1)
    T0 = rng.random(nx * ny) * 100; T = T0 + rng.random(nx * ny)
    hx, hy, dt = 100.0 / nx, 100.0 / ny, 1.0
    nz, r = O.assemble_heat(nx, ny, hx, hy, dt, T, T0, rowptr, colidx)
    A = to_scipy(nx * ny, 1, rowptr, colidx, nz)
    # linear problem: r(T + d) = r(T) + A d exactly up to rounding
    d = rng.random(nx * ny)
    _, r2 = O.assemble_heat(nx, ny, hx, hy, dt, T + d, T0, rowptr, colidx)
    assert np.allclose(r2, r + A @ d, rtol=1e-12, atol=1e-12)
    diag = 1.0 / dt + 2 / hx ** 2 + 2 / hy ** 2
    assert np.allclose(A.diagonal(), diag)


@pytest.mark.parametrize("dims", [(5, 4), (2, 3), (1, 1)])
def test_generic_cache_fill_reproduces_heat_assembly(O, dims):
    """fill_equation_entries! on the GenericAutoDiffCache of SimpleHeatEquation gives the Jacobian and residual of the
    dedicated heat assembly (two routes to the same LinearizedSystem; periodic aliasing on tiny grids merged)."""
    nx, ny = dims
    I, J = O.heat_pattern(nx, ny)
    rowptr, colidx = O.csr_from_coo(I, J, nx * ny)
    rng = np.random.default_rng(4)
    T = rng.uniform(0, 100, nx * ny); T0 = rng.uniform(0, 100, nx * ny)
    nz_ref, r_ref = O.assemble_heat(nx, ny, 0.7, 1.3, 0.5, T, T0, rowptr, colidx)
    vpos, dpos, pos, ent = O.generic_cache_heat(nx, ny, 0.7, 1.3, 0.5, T, T0, rowptr, colidx)
    nz = np.full(colidx.shape[0], np.nan); r = np.full(nx * ny, np.nan)
    O.generic_fill(1, 1, vpos, dpos, pos, ent, nz, r)
    assert np.allclose(nz, nz_ref, rtol=1e-13, atol=0) and np.allclose(r, r_ref, rtol=1e-12, atol=1e-12)
    # without diagonal positions the residual comes from the first slot of each entity: same value in every slot
    nz2 = np.zeros_like(nz); r2 = np.zeros_like(r)
    O.generic_fill(1, 1, vpos, None, pos, ent, nz2, r2)
    assert np.array_equal(nz2, nz) and np.array_equal(r2, r)


def test_load_balanced_endpoint_properties(O):
    """test/utils.jl:10-30, verbatim: endpoints, block widths floor/ceil(n/m), exact cover."""
    for n in range(1, 101):
        for m in range(1, n + 1):
            assert O.load_balanced_endpoint(0, n, m) == 0 and O.load_balanced_endpoint(m, n, m) == n
            count = 0
            for i in range(1, m + 1):
                start, stop = O.load_balanced_endpoint(i - 1, n, m), O.load_balanced_endpoint(i, n, m)
                assert start < stop
                assert stop - start in (n // m, -(-n // m))
                count += stop - start
            assert count == n
    assert O.load_balanced_interval(2, 10, 3) == (5, 7)
    # partition_linear (src/partitioning.jl:12-18) as shipped in the library: ceil(i / (n / m))
    assert O.partition_linear(3, 10).tolist() == [1, 1, 1, 2, 2, 2, 3, 3, 3, 3]


def test_scalar_and_multimodel_known_answers(O):
    """ScalarTestSystem: (X - X0)/dt - f = 0 from X0 = 0, f = 1, dt = 1 gives X = 1 (test/test_systems/scalar.jl:14-22);
    two of them with forces +1 / -1 and the skew-symmetric ScalarTestCrossTerm (X_T - X_S) give XA = 1/3, XB = -1/3
    (test/test_systems/multimodel.jl:4-37) — through the generic-cache fill and, for the grouped form, the Schur
    reduction with Jutul's sign convention dx = -(J \\ r)."""
    dt, X0 = 1.0, 0.0
    # single model: one entity, one slot, one equation, one partial
    vpos = np.array([1, 2]); pos = np.array([[1]]); dpos = np.array([1])
    ent = np.array([[[(X0 - X0) / dt - 1.0, 1.0 / dt]]])          # value after apply_forces! (diag_part -= force), partial
    nz = np.zeros(1); r = np.zeros(1)
    O.generic_fill(1, 1, vpos, dpos, pos, ent, nz, r)
    assert X0 - r[0] / nz[0] == 1.0
    # multimodel, groups = [1, 2]: B = dA/dXA, C = dA/dXB, D = dB/dXA, E = dB/dXB at X = 0
    XA = XB = 0.0
    rA = (XA - 0.0) / dt - 1.0 + (XA - XB); rB = (XB - 0.0) / dt + 1.0 - (XA - XB)
    B = np.array([[1.0 / dt + 1.0]]); C = [np.array([[-1.0]])]; D = [np.array([[-1.0]])]; E = [np.array([[1.0 / dt + 1.0]])]
    a = O.schur_prepare(np.array([rA]), C, E, [np.array([rB])])
    S = O.schur_mul(B, C, D, E, np.array([1.0]))
    dx_solver = a / S                                              # what the Krylov solver returns for the 1x1 operator
    x, y = O.schur_dx_update(D, E, [np.array([rB])], dx_solver)
    assert np.isclose(XA + x[0], 1.0 / 3.0, rtol=1e-15) and np.isclose(XB + y[0][0], -1.0 / 3.0, rtol=1e-15)
    # ungrouped (single sparse matrix): the same answer from the full 2x2 system
    Jm = np.array([[2.0, -1.0], [-1.0, 2.0]])
    assert np.allclose(-np.linalg.solve(Jm, [rA, rB]), [1.0 / 3.0, -1.0 / 3.0])


def test_block_spmv_reference_fixture_system_1(O):
    """system_1() of the reference's test/linalg/linear_operators.jl:60-78: A = sparse([1, 1, 2], [1, 2, 1], [a, b, c]) with 2x2
    SMatrix entries, X = [(pi, 3), (sqrt 3, 23.1)]. The property that file states (:119-123) — the block operator on the
    block-ordered vector equals the scalar matrix of block_system_to_scalar(reorder = true) (:8-37, entry[k, l] at
    ((k-1) n + i, (l-1) n + j)) on the equation-ordered vector, mapped back by equation_major_to_block_major_view — is checked
    here with the scalar matrix written out by hand, so it pins the oracle's column-major block convention to the reference's."""
    a = np.array([[1.0, 2], [3, 4]]); b = np.array([[5.0, 6], [7, 8]]); c = np.array([[9.0, 10], [11, 12]])
    x = np.array([[np.pi, 3.0], [np.sqrt(3.0), 23.1]])          # x[i] = block i
    # CSR of the 2x2 block matrix: row 1 = {a @ col 1, b @ col 2}, row 2 = {c @ col 1}; blocks stored column-major
    rowptr = np.array([1, 3, 4], dtype=np.int64); colidx = np.array([1, 2, 1], dtype=np.int64)
    nz = np.concatenate([m.flatten(order="F") for m in (a, b, c)])
    y = O.spmv(2, 2, rowptr, colidx, nz, x.ravel())
    # scalar matrix in equation-major order, by hand: rows/cols (k-1)*2 + i
    A_s = np.array([[1.0, 5, 2, 6],
                    [9.0, 0, 10, 0],
                    [3.0, 7, 4, 8],
                    [11.0, 0, 12, 0]])
    X_s = np.array([x[0, 0], x[1, 0], x[0, 1], x[1, 1]])
    res_s = A_s @ X_s
    renum = np.array([res_s[0], res_s[2], res_s[1], res_s[3]])   # equation_major_to_block_major_view(res_s, 2)
    assert np.allclose(y, renum, rtol=1e-15, atol=0)
    assert np.allclose(y, np.concatenate([a @ x[0] + b @ x[1], c @ x[0]]), rtol=1e-15, atol=0)
