"""The oracle (oracle/fe_oracle.py) pinned against the reference: frozen golden outputs always,
and the live unmodified reference when /root/reference is present (authoring container)."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import FULL_CASES, bc_sets_of
from oracle import fe_oracle as orc
from oracle import ref_runner


def rel_max(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


def test_tables_bit_exact(golden):
    z = golden("tables_and_single")
    for et, tag in ((orc.QUAD4, "4"), (orc.QUAD9, "9")):
        dH, H, w = orc.tables(et)
        assert np.array_equal(dH, z["dH" + tag])
        assert np.array_equal(H, z["H" + tag])
        assert np.array_equal(w, z["w" + tag])


def test_single_element_matrices(golden):
    z = golden("tables_and_single")
    ke4 = orc.element_matrices(np.arange(4)[None, :], z["X4"])[0]
    assert rel_max(ke4, z["ke4"]) < 1e-14
    assert abs(np.trace(ke4) - 7.238136187218806) < 1e-12
    ke9 = orc.element_matrices(z["conn9"] - 1, z["X9"])[0]
    assert rel_max(ke9, z["ke9"]) < 1e-13
    # unit-square Quad4 first row
    unit = np.array([[0, 0], [1, 0], [1, 1], [0, 1.0]])
    row = orc.element_matrices(np.arange(4)[None, :], unit)[0][0]
    assert np.allclose(row, [0.666637645, 0.25, -0.166637645, -0.25, -0.333362355, -0.25, -0.166637645, 0.25],
                       atol=1e-9)


@pytest.mark.parametrize("name", FULL_CASES)
def test_assembly_structure_and_values(golden, name):
    z = golden(name)
    conn0 = z["conn"] - 1
    K = orc.assemble_csr(conn0, z["coords"])
    assert K.indices.dtype == np.int32
    assert np.array_equal(K.indptr, z["k0_indptr"])
    assert np.array_equal(K.indices, z["k0_indices"])
    assert rel_max(K.data, z["k0_data"]) < 1e-12
    ke = orc.element_matrices(conn0[: len(z["ke"])], z["coords"])
    assert rel_max(ke, z["ke"]) < 1e-13
    # value-free pattern and the intermediate symbolic structures agree with it
    indptr, indices = orc.pattern_csr(conn0, z["coords"].shape[0])
    assert np.array_equal(indptr, z["k0_indptr"]) and np.array_equal(indices, z["k0_indices"])
    nptr, nadj = orc.node_adjacency(conn0, z["coords"].shape[0])
    assert 4 * nptr[-1] == K.nnz
    rows0 = indices[indptr[0]:indptr[1]]
    assert np.array_equal(rows0[0::2] // 2, nadj[nptr[0]:nptr[1]])
    eptr, einc = orc.node_incidence(conn0, z["coords"].shape[0])
    assert eptr[-1] == conn0.size
    assert np.all(conn0[einc // 16, einc % 16] == np.repeat(np.arange(len(eptr) - 1), np.diff(eptr)))


@pytest.mark.parametrize("name", FULL_CASES)
def test_bc_and_solution(golden, name):
    z = golden(name)
    conn0 = z["conn"] - 1
    d, nn = orc.boundary_dofs(bc_sets_of(z))
    assert np.array_equal(d, z["dofDir"]) and np.array_equal(nn, z["dofNeu"])
    K = orc.assemble_csr(conn0, z["coords"])
    Kbc, b = orc.apply_bc_reference(K, d, nn)
    assert np.array_equal(b, z["brhs"])
    # reference pattern ⊇ canonical; extras are stored zeros in Dirichlet rows
    Kref = sp.csr_matrix((z["k_data"], z["k_indices"], z["k_indptr"]), shape=K.shape)
    n = K.shape[0]
    ours = set(zip(np.repeat(np.arange(n), np.diff(Kbc.indptr)).tolist(), Kbc.indices.tolist()))
    rr = np.repeat(np.arange(n), np.diff(Kref.indptr))
    isdir = np.zeros(n, bool)
    isdir[d] = True
    extras = 0
    for r, c, v in zip(rr.tolist(), Kref.indices.tolist(), Kref.data.tolist()):
        if (r, c) not in ours:
            assert v == 0.0 and isdir[r]
            extras += 1
    assert len(ours) + extras == Kref.nnz
    assert rel_max(Kbc.toarray(), Kref.toarray()) < 1e-12
    U = orc.solve_direct(Kbc, b)
    assert np.linalg.norm(U - z["U"]) / np.linalg.norm(z["U"]) < 1e-10
    # the SPD lifted system + Jacobi-PCG (the GPU algorithm) reaches the same solution
    A, rhs = orc.lifted_system(K, d, b)
    assert abs(A - A.T).max() < 1e-12
    x, it, rr_ = orc.jacobi_pcg(A, rhs, rtol=1e-11)
    assert np.linalg.norm(x - z["U"]) / np.linalg.norm(z["U"]) < 1e-8


def test_scalar_known_answers(golden):
    from feprogram_b200 import mesh
    z = golden("scalars")
    conn, coords, bc = mesh.structured_quad4(64)
    K, Kbc, b, U = orc.solve_reference_path(conn - 1, coords, bc)
    g = z["q4_64"]
    assert K.nnz == int(g[0]) == 148996
    assert abs(np.sqrt((K.data ** 2).sum()) - g[1]) < 1e-10 * g[1]
    assert abs(K.diagonal().sum() - g[2]) < 1e-10 * g[2]
    assert abs(U.min() - g[6]) < 1e-7 and abs(U.max() - g[7]) < 1e-7
    assert abs(np.linalg.norm(U) - g[8]) < 1e-8 * g[8]


@pytest.mark.skipif(not ref_runner.available(), reason="reference tree not present on this box")
@pytest.mark.parametrize("kind,n", [(4, 9), (9, 5)])
def test_live_reference(kind, n):
    """Fresh meshes (not in the fixtures) straight against the unmodified reference."""
    from feprogram_b200 import mesh
    gen = mesh.structured_quad4 if kind == 4 else mesh.structured_quad9
    conn, coords, bc = gen(n)
    mx = n if kind == 4 else 2 * n
    coords = mesh.perturb_interior(coords, mx, mx, 0.25, seed=11)
    conn, coords, bc, _ = mesh.shuffle_numbering(conn, coords, bc, seed=5)
    Kref = ref_runner.assemble_prebc(conn, coords)
    K = orc.assemble_csr(conn - 1, coords)
    assert np.array_equal(K.indptr, Kref.indptr) and np.array_equal(K.indices, Kref.indices)
    assert rel_max(K.data, Kref.data) < 1e-12
    run = ref_runner.full_run(conn, coords, bc)
    _, Kbc, b, U = orc.solve_reference_path(conn - 1, coords, bc)
    assert np.array_equal(b, run["brhs"])
    assert np.linalg.norm(U - run["U"]) / np.linalg.norm(run["U"]) < 1e-10
    d4, h4, w4 = ref_runner.tables(kind)
    t = orc.tables(kind)
    assert np.array_equal(t[0], d4) and np.array_equal(t[1], h4) and np.array_equal(t[2], w4)


def test_mass_oracle_against_reference_built_golden():
    """orc.mass_matrices vs N^T N assembled from the reference's own H / gpWei / Hrs objects (frozen)."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "mass_single.npz"))
    assert np.abs(orc.mass_matrices(np.arange(4)[None, :], g["X4"])[0] - g["me4"]).max() < 1e-14 * np.abs(g["me4"]).max()
    assert np.abs(orc.mass_matrices(g["conn9"] - 1, g["X9"])[0] - g["me9"]).max() < 1e-14 * np.abs(g["me9"]).max()
    assert np.abs(orc.mass_matrices(g["conn"] - 1, g["coords"]) - g["me"]).max() < 1e-14 * np.abs(g["me"]).max()
    M = orc.assemble_csr(g["conn"] - 1, g["coords"], mass=True)
    K = orc.assemble_csr(g["conn"] - 1, g["coords"])
    assert np.array_equal(M.indptr, K.indptr) and np.array_equal(M.indices, K.indices)   # one pattern for K and M


@pytest.mark.skipif(not ref_runner.available(), reason="reference tree not present on this box")
def test_mass_live_reference_objects():
    from feprogram_b200 import mesh
    conn, coords, _ = mesh.structured_quad9(3)
    coords = mesh.perturb_interior(coords, 6, 6, 0.2, seed=11)
    fem = ref_runner.make_problem(conn, coords)
    want = np.array([ref_runner.mass_matrix(fem, conn[e] - 1) for e in range(len(conn))])
    got = orc.mass_matrices(conn - 1, coords)
    assert np.abs(got - want).max() < 1e-14 * np.abs(want).max()


def test_gradient_oracle_against_reference_getHgrad_golden():
    """orc.gauss_gradients vs the reference's own Elem.getHgrad applied at every Gauss point (frozen)."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "post_gradients.npz"))
    for k in (4, 9):
        got = orc.gauss_gradients(g["conn%d" % k] - 1, g["coords%d" % k], g["U%d" % k])
        assert np.abs(got - g["grad%d" % k]).max() < 1e-13 * np.abs(got).max()
    # a linear field has a constant gradient; the nodal strain recovers it exactly
    from feprogram_b200 import mesh
    conn, coords, _ = mesh.structured_quad4(6, 5)
    coords = mesh.perturb_interior(coords, 6, 5, 0.2, seed=1)
    A = np.array([[0.3, -1.2], [2.0, 0.7]])
    U = (coords @ A.T).ravel()
    eps = orc.nodal_strain(conn - 1, coords, U)
    assert np.abs(eps - np.array([A[0, 0], A[1, 1], A[0, 1] + A[1, 0]])[None, :]).max() < 1e-12


@pytest.mark.skipif(not ref_runner.available(), reason="reference tree not present on this box")
def test_gradient_live_reference():
    from feprogram_b200 import mesh
    conn, coords, _ = mesh.structured_quad9(2, 3)
    coords = mesh.perturb_interior(coords, 4, 6, 0.2, seed=4)
    U = np.random.default_rng(5).standard_normal(2 * coords.shape[0])
    fem = ref_runner.make_problem(conn, coords)
    want = ref_runner.gauss_gradients(fem, conn, U)
    got = orc.gauss_gradients(conn - 1, coords, U)
    assert np.abs(got - want).max() < 1e-13 * np.abs(want).max()


@pytest.mark.parametrize("name", FULL_CASES)
def test_poisson_submatrix_of_the_reference_matrix(golden, name):
    """FE_OP_POISSON's oracle is K[0::2, 0::2] of the REFERENCE's own pre-BC matrix (frozen in tests/golden), so the
    scalar operator is pinned to the reference too; its pattern is the node adjacency and it annihilates constants."""
    z = golden(name)
    n2 = len(z["k0_indptr"]) - 1
    Kref = sp.csr_matrix((z["k0_data"], z["k0_indices"], z["k0_indptr"]), shape=(n2, n2))
    Ks_ref = orc.poisson_submatrix(Kref)
    conn0 = z["conn"] - 1
    Ks = orc.poisson_submatrix(orc.assemble_csr(conn0, z["coords"]))
    nptr, nadj = orc.node_adjacency(conn0, z["coords"].shape[0])
    assert np.array_equal(Ks.indptr, nptr) and np.array_equal(Ks.indices, nadj)
    assert np.array_equal(Ks_ref.indptr, Ks.indptr) and np.array_equal(Ks_ref.indices, Ks.indices)
    assert rel_max(Ks.data, Ks_ref.data) < 1e-12
    odd = Kref[1::2, :][:, 1::2].tocsr()                 # odd/odd block: the same sums in another order
    odd.sort_indices()
    assert rel_max(odd.data, Ks_ref.data) < 1e-13
    assert np.abs(Ks_ref @ np.ones(Ks.shape[0])).max() < 1e-12 * np.abs(Ks.data).max() * 10
    assert abs(Ks_ref - Ks_ref.T).max() < 1e-13 * np.abs(Ks.data).max()
    # whole scalar path: Dirichlet rows are unit rows, loaded nodes carry 1.0, the solve satisfies the system
    Ks2, Kbc, b, u = orc.solve_poisson_path(conn0, z["coords"], bc_sets_of(z))
    d, nn = orc.poisson_boundary_dofs(bc_sets_of(z))
    assert np.array_equal(b[nn], np.ones(len(nn))) and np.count_nonzero(b) == len(nn)
    assert np.abs(Kbc @ u - b).max() < 1e-9 * max(1.0, np.abs(u).max())
