"""Parity of the CUDA path (through the C ABI / torch custom ops / FemProblem facade) with the oracle
and with the frozen outputs of the unmodified reference (tests/golden).  Tolerances (north_star):
CSR structure bit-exact; values <= 1e-12 relative (norm-wise); solution <= 1e-8 rel-L2."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import FULL_CASES, bc_sets_of
from oracle import fe_oracle as orc

pytestmark = pytest.mark.gpu

VAL_TOL = 1e-12
SOL_TOL = 1e-8


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    from feprogram_b200 import build
    build.build()
    from feprogram_b200 import ops  # noqa: F401
    assert torch.cuda.is_available()
    return torch


def rel_max(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / np.abs(np.asarray(b)).max()


def dev_mesh(torch, conn1, coords):
    conn0 = torch.from_numpy((np.asarray(conn1, dtype=np.int64) - 1).astype(np.int32)).cuda()
    xy = torch.from_numpy(np.ascontiguousarray(coords, dtype=np.float64)).cuda()
    return conn0, xy


def test_single_element_matrices_against_reference(torch_cuda, golden):
    torch = torch_cuda
    z = golden("tables_and_single")
    c4, x4 = dev_mesh(torch, np.arange(1, 5)[None, :], z["X4"])
    ke4 = torch.ops.feprogram.element_matrices(c4, x4, 4, 0).cpu().numpy()[0]
    assert rel_max(ke4, z["ke4"]) < 1e-13
    c9, x9 = dev_mesh(torch, z["conn9"], z["X9"])
    ke9 = torch.ops.feprogram.element_matrices(c9, x9, 9, 0).cpu().numpy()[0]
    assert rel_max(ke9, z["ke9"]) < 1e-13
    assert abs(np.trace(ke4) - 7.238136187218806) < 1e-12


@pytest.mark.parametrize("name", FULL_CASES)
def test_element_matrices_batch(torch_cuda, golden, name):
    torch = torch_cuda
    z = golden(name)
    conn0, xy = dev_mesh(torch, z["conn"], z["coords"])
    et = z["conn"].shape[1]
    ke = torch.ops.feprogram.element_matrices(conn0, xy, et, 0).cpu().numpy()
    assert rel_max(ke[: len(z["ke"])], z["ke"]) < 1e-13                       # the reference's getKe
    assert rel_max(ke, orc.element_matrices(z["conn"] - 1, z["coords"])) < 1e-13  # the oracle, all elements
    kw = torch.ops.feprogram.element_matrices(conn0, xy, et, 1).cpu().numpy()
    assert rel_max(kw, orc.element_matrices(z["conn"] - 1, z["coords"], weighted=True)) < 1e-13


@pytest.mark.parametrize("name", FULL_CASES)
def test_symbolic_structures_bit_exact(torch_cuda, golden, name):
    torch = torch_cuda
    from feprogram_b200 import ops
    z = golden(name)
    conn0, _ = dev_mesh(torch, z["conn"], z["coords"])
    nnode = z["coords"].shape[0]
    pat = ops.symbolic(conn0, nnode, z["conn"].shape[1], 2)
    eptr, einc = orc.node_incidence(z["conn"] - 1, nnode)
    assert np.array_equal(pat.eptr.cpu().numpy(), eptr)
    assert np.array_equal(pat.einc.cpu().numpy(), einc)
    nptr, nadj = orc.node_adjacency(z["conn"] - 1, nnode)
    assert np.array_equal(pat.nptr.cpu().numpy(), nptr)
    assert np.array_equal(pat.nadj.cpu().numpy(), nadj)
    # the reference's own CSR structure, bit for bit (int32 like scipy's)
    assert pat.row_ptr.dtype == torch.int32 and pat.col_idx.dtype == torch.int32
    assert np.array_equal(pat.row_ptr.cpu().numpy(), z["k0_indptr"])
    assert np.array_equal(pat.col_idx.cpu().numpy(), z["k0_indices"])
    # slot map: position of every element node inside the owner's adjacency list
    epos = pat.epos.cpu().numpy().reshape(-1, z["conn"].shape[1])
    conn0h = z["conn"] - 1
    owner = np.repeat(np.arange(nnode), np.diff(eptr))
    for k in range(0, len(einc), max(1, len(einc) // 97)):
        e = einc[k] // 16
        seg = nadj[nptr[owner[k]]:nptr[owner[k] + 1]]
        assert np.array_equal(seg[epos[k]], conn0h[e])
    # 64-bit row pointers are the same numbers
    rp64, ci64 = torch.ops.feprogram.csr_expand(pat.nptr, pat.nadj, 2, True)
    assert rp64.dtype == torch.int64 and np.array_equal(rp64.cpu().numpy(), z["k0_indptr"])
    assert np.array_equal(ci64.cpu().numpy(), z["k0_indices"])


def run_facade(z, **kw):
    from feprogram_b200.elements import FemProblem
    fem = FemProblem(z["conn"], z["coords"], **kw)
    fem.setMatrix()
    fem.setBoundaryConditions(bc_sets_of(z), verbose=False)
    fem.assemble()
    return fem


@pytest.mark.parametrize("name", FULL_CASES)
def test_facade_assembly_bc_and_solution_against_reference(torch_cuda, golden, name):
    z = golden(name)
    fem = run_facade(z)
    K0 = fem.csr_prebc()
    assert K0.indices.dtype == np.int32
    assert np.array_equal(K0.indptr, z["k0_indptr"]) and np.array_equal(K0.indices, z["k0_indices"])
    assert rel_max(K0.data, z["k0_data"]) < VAL_TOL
    # post-BC system in the reference's own types
    K = fem.K
    assert sp.issparse(K) and K.format == "csc" and K.shape == (2 * fem.nnode, 2 * fem.nnode)
    b = np.asarray(fem.brhs.todense()).ravel()
    assert fem.brhs.format == "csc" and np.array_equal(b, z["brhs"])
    Kref = sp.csr_matrix((z["k_data"], z["k_indices"], z["k_indptr"]), shape=K.shape)
    assert rel_max(K.toarray(), Kref.toarray()) < VAL_TOL
    # reference pattern ⊇ ours, extras are stored zeros in Dirichlet rows
    Kc = K.tocsr()
    Kc.sort_indices()
    ours = set(zip(np.repeat(np.arange(K.shape[0]), np.diff(Kc.indptr)).tolist(), Kc.indices.tolist()))
    isdir = np.zeros(K.shape[0], bool)
    isdir[z["dofDir"]] = True
    rr = np.repeat(np.arange(K.shape[0]), np.diff(Kref.indptr))
    extras = [(r, c, v) for r, c, v in zip(rr.tolist(), Kref.indices.tolist(), Kref.data.tolist())
              if (r, c) not in ours]
    assert all(v == 0.0 and isdir[r] for r, c, v in extras)
    assert len(ours) + len(extras) == Kref.nnz
    # solution
    U = fem.solveProblem()
    assert U.shape == (2 * fem.nnode,) and U.dtype == np.float64
    assert fem.solver_info["converged"]
    assert np.linalg.norm(U - z["U"]) / np.linalg.norm(z["U"]) < SOL_TOL
    assert np.array_equal(U[z["dofDir"]], b[z["dofDir"]])          # Dirichlet values exact (corner dofs = 1)


def test_getke_and_weighted_operator(torch_cuda, golden):
    z = golden("q4_16x16_pert")
    fem = run_facade(z)
    for e in range(3):
        ke = fem.getKe(z["conn"][e] - 1)
        assert isinstance(ke, np.matrix) and ke.shape == (8, 8)
        assert rel_max(ke, z["ke"][e]) < 1e-13
    z9 = golden("q9_6x6_pert_shuf")
    femw = run_facade(z9, operator="weighted")
    Kw = orc.assemble_csr(z9["conn"] - 1, z9["coords"], weighted=True)
    assert rel_max(femw.csr_prebc().data, Kw.data) < VAL_TOL


def test_determinism_bit_identical_reruns(torch_cuda, golden):
    torch = torch_cuda
    z = golden("q9_6x6_pert_shuf")
    fem = run_facade(z)
    v1 = fem._dev["vals"].clone()
    u1 = fem.solveProblem()
    fem.setMatrix()
    fem.assemble()
    assert torch.equal(v1, fem._dev["vals"])
    assert np.array_equal(u1, fem.solveProblem())
    from feprogram_b200 import mesh
    conn, coords, bc = mesh.structured_quad4(96)
    coords = mesh.perturb_interior(coords, 96, 96)
    z2 = dict(conn=conn, coords=coords, **{"bc%d" % i: np.array(sorted(s)) for i, s in enumerate(bc)})
    f1, f2 = run_facade(z2), run_facade(z2)
    assert torch.equal(f1._dev["vals"], f2._dev["vals"])


def test_spmv_and_bc_modes_against_scipy(torch_cuda, golden):
    torch = torch_cuda
    z = golden("q4_12x12_pert_shuf")
    fem = run_facade(z)
    st, pat = fem._dev, fem._dev["pattern"]
    K0 = fem.csr_prebc()
    rng = np.random.default_rng(3)
    x = rng.standard_normal(pat.nrows)
    xd = torch.from_numpy(x).cuda()
    y = torch.empty_like(xd)
    torch.ops.feprogram.spmv(pat.row_ptr, pat.col_idx, st["vals"], xd, y, None)
    assert rel_max(y.cpu().numpy(), K0 @ x) < 1e-13
    torch.ops.feprogram.spmv(pat.row_ptr, pat.col_idx, st["vals"], xd, y, st["dirmask"])
    ref = K0 @ x
    ref[z["dofDir"]] = 0.0
    assert rel_max(y.cpu().numpy(), ref) < 1e-13
    # node-block form of the same product (with and without the Dirichlet row mask)
    yn = torch.empty_like(xd)
    torch.ops.feprogram.spmv_nodal(pat.nptr, pat.nadj, st["vals"], xd, yn, None)
    assert rel_max(yn.cpu().numpy(), K0 @ x) < 1e-13
    torch.ops.feprogram.spmv_nodal(pat.nptr, pat.nadj, st["vals"], xd, yn, st["dirmask"])
    assert rel_max(yn.cpu().numpy(), ref) < 1e-13
    d = torch.ops.feprogram.csr_diagonal(pat.row_ptr, pat.col_idx, st["vals"]).cpu().numpy()
    assert np.array_equal(d, K0.diagonal())
    # symmetric lift in place == oracle's lifted system
    vals = st["vals"].clone()
    rhs = st["rhs"].clone()
    torch.ops.feprogram.apply_bc(1, pat.row_ptr, pat.col_idx, vals, rhs, st["dirmask"])
    A, b = orc.lifted_system(K0, z["dofDir"], z["brhs"])
    assert rel_max(vals.cpu().numpy(), A.data) < 1e-14
    assert rel_max(rhs.cpu().numpy(), b) < 1e-13
    # plain SPD solve on the lifted matrix (no mask) gives the same solution
    xs = torch.empty_like(rhs)
    info = torch.ops.feprogram.pcg(pat.row_ptr, pat.col_idx, vals, rhs, None, xs, 1e-11, 0.0, 100000, 25)
    assert info[3] == 1.0
    assert np.linalg.norm(xs.cpu().numpy() - z["U"]) / np.linalg.norm(z["U"]) < SOL_TOL


def test_cfg1_quad4_64_known_answers(torch_cuda, golden):
    """BASELINE config 1 (2-D Quad4 64x64): scalars frozen from the reference."""
    from feprogram_b200 import mesh
    g = golden("scalars")["q4_64"]
    conn, coords, bc = mesh.structured_quad4(64)
    z = dict(conn=conn, coords=coords, **{"bc%d" % i: np.array(sorted(s)) for i, s in enumerate(bc)})
    fem = run_facade(z)
    K0 = fem.csr_prebc()
    assert K0.nnz == 148996 == int(g[0])
    assert abs(np.sqrt((K0.data ** 2).sum()) - g[1]) < 1e-11 * g[1]
    assert abs(K0.diagonal().sum() - g[2]) < 1e-11 * g[2]
    U = fem.solveProblem()
    assert abs(U.min() - g[6]) < 1e-6 and abs(U.max() - g[7]) < 1e-6
    assert abs(np.linalg.norm(U) - g[8]) < 1e-8 * g[8]
    Ko, Kbc, b, Uo = orc.solve_reference_path(conn - 1, coords, bc)
    assert np.linalg.norm(U - Uo) / np.linalg.norm(Uo) < SOL_TOL


@pytest.mark.parametrize("kind,n", [(4, 300), (9, 120)])
def test_medium_mesh_against_oracle(torch_cuda, kind, n):
    """Perturbed geometry + shuffled numbering at a size only the vectorised oracle reaches quickly."""
    from feprogram_b200 import mesh
    gen = mesh.structured_quad4 if kind == 4 else mesh.structured_quad9
    conn, coords, bc = gen(n)
    mx = n if kind == 4 else 2 * n
    coords = mesh.perturb_interior(coords, mx, mx, 0.2, seed=0)
    conn, coords, bc, _ = mesh.shuffle_numbering(conn, coords, bc, seed=1)
    z = dict(conn=conn, coords=coords, **{"bc%d" % i: np.array(sorted(s)) for i, s in enumerate(bc)})
    fem = run_facade(z)
    K0 = fem.csr_prebc()
    Ko = orc.assemble_csr(conn - 1, coords)
    assert np.array_equal(K0.indptr, Ko.indptr) and np.array_equal(K0.indices, Ko.indices)
    assert rel_max(K0.data, Ko.data) < VAL_TOL
    fem.rtol = 1e-11
    U = fem.solveProblem()
    d, nn = orc.boundary_dofs(bc)
    Kbc, b = orc.apply_bc_reference(Ko, d, nn)
    Uo = orc.solve_direct(Kbc, b)
    assert np.linalg.norm(U - Uo) / np.linalg.norm(Uo) < SOL_TOL


def test_empty_and_degenerate_inputs(torch_cuda):
    torch = torch_cuda
    from feprogram_b200 import ops
    # mesh with an isolated (unused) node: its rows are empty, like the reference's
    conn = np.array([[1, 2, 4, 3]])
    coords = np.array([[0, 0], [1, 0], [0, 1], [1, 1], [5, 5.0]])
    conn0, xy = dev_mesh(torch, conn, coords)
    pat = ops.symbolic(conn0, 5, 4, 2)
    rp = pat.row_ptr.cpu().numpy()
    assert rp[-1] == 64 and rp[8] == rp[9] == rp[10] == 64
    vals = torch.empty(pat.nnz, dtype=torch.float64, device="cuda")
    torch.ops.feprogram.assemble(conn0, xy, pat.eptr, pat.einc, pat.epos, pat.nptr, pat.max_deg, vals, 4, 0)
    K = sp.csr_matrix((vals.cpu().numpy(), pat.col_idx.cpu().numpy(), rp), shape=(10, 10))
    Ko = orc.assemble_csr(conn - 1, coords)
    assert rel_max(K.toarray(), Ko.toarray()) < VAL_TOL
    # zero elements
    ke = torch.ops.feprogram.element_matrices(conn0[:0].contiguous(), xy, 4, 0)
    assert ke.shape == (0, 8, 8)
    # unsupported operator code is reported, not ignored
    from feprogram_b200._lib import FeError
    with pytest.raises(FeError):
        torch.ops.feprogram.element_matrices(conn0, xy, 4, 7)


def test_full_size_cfg2_properties(torch_cuda):
    """BASELINE config 2 (Quad4 4096x4096, 16.8 M elements): size-independent properties of K --
    closed-form nnz, rigid-body null space (2 translations + rotation), symmetry via SpMV, exact
    agreement of a window of rows with the oracle on the sub-mesh that determines them."""
    torch = torch_cuda
    from feprogram_b200 import mesh, ops
    n = 4096
    conn, coords, _ = mesh.structured_quad4(n, dtype=np.int32)
    conn0 = torch.from_numpy(conn - 1).cuda()
    xy = torch.from_numpy(coords).cuda()
    nnode = coords.shape[0]
    pat = ops.symbolic(conn0, nnode, 4, 2)
    assert pat.nnz == 4 * (3 * n + 1) ** 2 == 604078084
    vals = torch.empty(pat.nnz, dtype=torch.float64, device="cuda")
    torch.ops.feprogram.assemble(conn0, xy, pat.eptr, pat.einc, pat.epos, pat.nptr, pat.max_deg, vals, 4, 0)
    y = torch.empty(pat.nrows, dtype=torch.float64, device="cuda")
    scale = float(vals.abs().max())
    modes = []
    tx = torch.zeros(pat.nrows, dtype=torch.float64, device="cuda"); tx[0::2] = 1.0
    ty = torch.zeros(pat.nrows, dtype=torch.float64, device="cuda"); ty[1::2] = 1.0
    rot = torch.empty(pat.nrows, dtype=torch.float64, device="cuda"); rot[0::2] = -xy[:, 1]; rot[1::2] = xy[:, 0]
    for m in (tx, ty, rot):
        torch.ops.feprogram.spmv(pat.row_ptr, pat.col_idx, vals, m, y, None)
        assert float(y.abs().max()) < 1e-12 * scale * 10
        modes.append(m)
    g = torch.Generator(device="cuda").manual_seed(0)
    a = torch.randn(pat.nrows, dtype=torch.float64, device="cuda", generator=g)
    b = torch.randn(pat.nrows, dtype=torch.float64, device="cuda", generator=g)
    Ka, Kb = torch.empty_like(a), torch.empty_like(b)
    torch.ops.feprogram.spmv(pat.row_ptr, pat.col_idx, vals, a, Ka, None)
    torch.ops.feprogram.spmv(pat.row_ptr, pat.col_idx, vals, b, Kb, None)
    lhs, rhs_ = float(b @ Ka), float(a @ Kb)
    assert abs(lhs - rhs_) < 1e-10 * max(abs(lhs), 1.0)
    # rows of the nodes in mesh rows 1000..1001 depend only on element rows 999..1001: oracle on that strip
    j0 = 999
    sub_conn, sub_coords, _ = mesh.structured_quad4(n, 3)
    sub_coords = sub_coords.copy()
    sub_coords[:, 1] = np.linspace(0, 1, n + 1)[j0 + (np.arange(sub_coords.shape[0]) // (n + 1))]
    Ko = orc.assemble_csr(sub_conn - 1, sub_coords)
    w = n + 1
    r0, r1 = 2 * (1000 * w), 2 * (1002 * w)              # global rows of mesh rows 1000, 1001
    rp = pat.row_ptr[r0:r1 + 1].cpu().numpy().astype(np.int64)
    got = vals[rp[0]:rp[-1]].cpu().numpy()
    s0, s1 = 2 * (1 * w), 2 * (3 * w)                    # the same nodes inside the strip
    want = Ko.data[Ko.indptr[s0]:Ko.indptr[s1]]
    assert got.shape == want.shape and rel_max(got, want) < VAL_TOL
    cols = pat.col_idx[rp[0]:rp[-1]].cpu().numpy().astype(np.int64)
    assert np.array_equal(cols - 2 * (j0 * w), Ko.indices[Ko.indptr[s0]:Ko.indptr[s1]])


def test_tile_plan_exists_for_regular_meshes(torch_cuda):
    """Power-of-two structured meshes (the benchmark family) must get the tiled kernel, not the fallback."""
    torch = torch_cuda
    from feprogram_b200 import mesh, ops
    for kind, gen, n in ((4, mesh.structured_quad4, 128), (3, mesh.structured_tri3, 64)):
        conn, coords, _ = gen(n)
        conn0, xy = dev_mesh(torch, conn, coords)
        pat = ops.symbolic(conn0, coords.shape[0], kind, 2)
        plan = ops.tile_plan(conn0, xy, pat)
        assert plan is not None, (kind, n)
        v1 = torch.empty(pat.nnz, dtype=torch.float64, device="cuda")
        v2 = torch.full((pat.nnz,), float("nan"), dtype=torch.float64, device="cuda")
        ops.assemble_values(conn0, xy, pat, None, v1)
        ops.assemble_values(conn0, xy, pat, plan, v2)
        assert torch.equal(v1, v2)


@pytest.mark.parametrize("kind,n,shuffle", [(4, 70, False), (4, 45, True), (3, 40, False), (3, 33, True)])
def test_tiled_and_row_kernels_agree_bitwise(torch_cuda, kind, n, shuffle):
    """The tiled kernel (Morton tiles and input-order tiles) and the row-tile kernel produce the same
    bits: same per-pair arithmetic, same ascending-element summation order; and K is bit-wise symmetric."""
    torch = torch_cuda
    from feprogram_b200 import mesh, ops
    gen = mesh.structured_quad4 if kind == 4 else mesh.structured_tri3
    conn, coords, bc = gen(n, n + 3)
    coords = mesh.perturb_interior(coords, n, n + 3, 0.2, seed=2)
    if shuffle:
        conn, coords, bc, _ = mesh.shuffle_numbering(conn, coords, bc, seed=4)
    conn0, xy = dev_mesh(torch, conn, coords)
    pat = ops.symbolic(conn0, coords.shape[0], kind, 2)
    v_rows = torch.empty(pat.nnz, dtype=torch.float64, device="cuda")
    ops.assemble_values(conn0, xy, pat, None, v_rows)
    checked = 0
    for morton in (True, False):
        for kernel in ("ws", "phased"):    # warp-specialised (mbarrier / TMA) kernel and barrier-phased kernel
            plan = ops.tile_plan(conn0, xy, pat, morton=morton, kernel=kernel)
            if plan is None:   # tile capacities exceeded (e.g. input-order strips): the facade falls back to rows
                continue
            assert plan.warp_specialised == (kernel == "ws")
            checked += 1
            v = torch.full((pat.nnz,), float("nan"), dtype=torch.float64, device="cuda")
            ops.assemble_values(conn0, xy, pat, plan, v)
            assert torch.equal(v, v_rows), (kernel, morton)
    assert checked >= 2
    K = sp.csr_matrix((v_rows.cpu().numpy(), pat.col_idx.cpu().numpy(), pat.row_ptr.cpu().numpy()),
                      shape=(pat.nrows, pat.nrows))
    assert abs(K - K.T).max() == 0.0
    Ko = orc.assemble_csr(conn - 1, coords, kind)
    assert np.array_equal(K.indptr, Ko.indptr) and np.array_equal(K.indices, Ko.indices)
    assert rel_max(K.data, Ko.data) < VAL_TOL


@pytest.mark.parametrize("n,perturb,shuffle", [(16, False, False), (21, True, False), (13, True, True)])
def test_quad9_tiled_kernel_matches_row_kernel_bitwise(torch_cuda, n, perturb, shuffle):
    """Quad9 in the tiled kernel (five threads per element) == row-tile kernel, bit for bit, == oracle at 1e-12."""
    torch = torch_cuda
    from feprogram_b200 import mesh, ops
    conn, coords, bc = mesh.structured_quad9(n, n + 2)
    if perturb:
        coords = mesh.perturb_interior(coords, 2 * n, 2 * (n + 2), 0.15, seed=9)
    if shuffle:
        conn, coords, bc, _ = mesh.shuffle_numbering(conn, coords, bc, seed=10)
    conn0, xy = dev_mesh(torch, conn, coords)
    pat = ops.symbolic(conn0, coords.shape[0], 9, 2)
    v_rows = torch.empty(pat.nnz, dtype=torch.float64, device="cuda")
    ops.assemble_values(conn0, xy, pat, None, v_rows)
    plan = ops.tile_plan(conn0, xy, pat)
    assert plan is not None or shuffle or perturb
    if plan is not None:
        v = torch.full((pat.nnz,), float("nan"), dtype=torch.float64, device="cuda")
        ops.assemble_values(conn0, xy, pat, plan, v)
        assert torch.equal(v, v_rows)
    K = sp.csr_matrix((v_rows.cpu().numpy(), pat.col_idx.cpu().numpy(), pat.row_ptr.cpu().numpy()),
                      shape=(pat.nrows, pat.nrows))
    Ko = orc.assemble_csr(conn - 1, coords, 9)
    assert np.array_equal(K.indptr, Ko.indptr) and np.array_equal(K.indices, Ko.indices)
    assert rel_max(K.data, Ko.data) < VAL_TOL
    assert abs(K - K.T).max() == 0.0


def test_gmsh_style_inputs_and_int64_row_pointers(torch_cuda, golden):
    """(1) The reference's real input types: connectivity as a list of uint64 arrays, coordinates (nnode,3).
    (2) The 64-bit row-pointer code path (used when nnz >= 2^31) gives the same SpMV / PCG results."""
    torch = torch_cuda
    from feprogram_b200.elements import FemProblem
    z = golden("q4_16x16_pert")
    conn_list = [np.asarray(r, dtype=np.uint64) for r in z["conn"]]
    xyz = np.concatenate([z["coords"], np.zeros((z["coords"].shape[0], 1))], axis=1)
    fem = FemProblem(conn_list, xyz)
    fem.setMatrix()
    fem.setBoundaryConditions(bc_sets_of(z), verbose=False)
    fem.assemble()
    K0 = fem.csr_prebc()
    assert np.array_equal(K0.indptr, z["k0_indptr"]) and np.array_equal(K0.indices, z["k0_indices"])
    assert rel_max(K0.data, z["k0_data"]) < VAL_TOL
    U = fem.solveProblem()
    assert np.linalg.norm(U - z["U"]) / np.linalg.norm(z["U"]) < SOL_TOL
    st, pat = fem._dev, fem._dev["pattern"]
    rp64, ci = torch.ops.feprogram.csr_expand(pat.nptr, pat.nadj, 2, True)
    x = torch.randn(pat.nrows, dtype=torch.float64, device="cuda")
    y32, y64 = torch.empty_like(x), torch.empty_like(x)
    torch.ops.feprogram.spmv(pat.row_ptr, pat.col_idx, st["vals"], x, y32, None)
    torch.ops.feprogram.spmv(rp64, ci, st["vals"], x, y64, None)
    assert torch.equal(y32, y64)
    s32, s64 = torch.empty_like(x), torch.empty_like(x)
    i32 = torch.ops.feprogram.pcg(pat.row_ptr, pat.col_idx, st["vals"], st["rhs"], st["dirmask"], s32, 1e-10, 0.0, 5000, 20)
    i64 = torch.ops.feprogram.pcg(rp64, ci, st["vals"], st["rhs"], st["dirmask"], s64, 1e-10, 0.0, 5000, 20)
    assert torch.equal(s32, s64) and i32[0] == i64[0] and i32[3] == 1.0
    # CSR-walk PCG and node-block PCG reach the same solution (different summation grouping, same tolerance)
    sn = torch.empty_like(x)
    torch.ops.feprogram.pcg(pat.row_ptr, pat.col_idx, st["vals"], st["rhs"], st["dirmask"], sn, 1e-10, 0.0, 5000, 20,
                            pat.nptr, pat.nadj)
    assert np.linalg.norm(sn.cpu().numpy() - z["U"]) / np.linalg.norm(z["U"]) < SOL_TOL
    assert np.linalg.norm(s32.cpu().numpy() - z["U"]) / np.linalg.norm(z["U"]) < SOL_TOL


def test_maxit_and_nonconvergence_are_reported(torch_cuda, golden):
    z = golden("q9_4x4")
    fem = run_facade(z)
    fem.maxit = 3
    with pytest.warns(RuntimeWarning, match="iteration limit"):
        U = fem.solveProblem()
    assert fem.solver_info["iterations"] == 3 and not fem.solver_info["converged"] and not fem.solver_info["breakdown"]
    assert np.all(np.isfinite(U))


@pytest.mark.parametrize("kind,n", [(4, 150), (3, 90), (9, 40)])
def test_kernels_stay_inside_their_output_buffers(torch_cuda, kind, n):
    """compute-sanitizer is not available on the GPU pool, so out-of-bounds stores are caught with guard
    zones: the value array sits inside a larger buffer filled with a sentinel; both assembly kernels, SpMV
    and PCG must leave the guards untouched and write every entry in between."""
    torch = torch_cuda
    from feprogram_b200 import mesh, ops
    gen = {4: mesh.structured_quad4, 3: mesh.structured_tri3, 9: mesh.structured_quad9}[kind]
    conn, coords, bc = gen(n, n + 1)
    mx = 2 * n if kind == 9 else n
    coords = mesh.perturb_interior(coords, mx, mx + (2 if kind == 9 else 1), 0.15, seed=12)
    conn0, xy = dev_mesh(torch, conn, coords)
    pat = ops.symbolic(conn0, coords.shape[0], kind, 2)
    plan = ops.tile_plan(conn0, xy, pat)
    G, sentinel = 4096, -7.25e300

    def guarded(n_):
        big = torch.full((n_ + 2 * G,), sentinel, dtype=torch.float64, device="cuda")
        return big, big[G:G + n_]

    for p in ([plan, None] if plan is not None else [None]):
        big, vals = guarded(pat.nnz)
        ops.assemble_values(conn0, xy, pat, p, vals)
        torch.cuda.synchronize()
        assert bool((big[:G] == sentinel).all()) and bool((big[-G:] == sentinel).all())
        assert not bool((vals == sentinel).any())
    x = torch.randn(pat.nrows, dtype=torch.float64, device="cuda")
    for op in ("csr", "nodal"):
        big, y = guarded(pat.nrows)
        if op == "csr":
            torch.ops.feprogram.spmv(pat.row_ptr, pat.col_idx, vals, x, y, None)
        else:
            torch.ops.feprogram.spmv_nodal(pat.nptr, pat.nadj, vals, x, y, None)
        torch.cuda.synchronize()
        assert bool((big[:G] == sentinel).all()) and bool((big[-G:] == sentinel).all())
        assert not bool((y == sentinel).any())


# ---------------------------------------------------------------------------------------------------
# mass operator (FE_OP_MASS): N^T N on the reference's H / gpWei tables (extension)
# ---------------------------------------------------------------------------------------------------
def test_mass_element_matrices_against_reference_tables(torch_cuda):
    """fe_element_matrices(op=mass) vs the product assembled from the reference's own fem.H / fem.gpWei /
    fem.Hrs objects (tests/golden/mass_single.npz, oracle/make_golden.py mass_golden)."""
    torch = torch_cuda
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "mass_single.npz"))
    for conn1, X, want in ((np.array([[1, 2, 3, 4]]), g["X4"], g["me4"][None]), (g["conn9"], g["X9"], g["me9"][None]),
                           (g["conn"], g["coords"], g["me"])):
        conn0, xy = dev_mesh(torch, conn1, X)
        me = torch.ops.feprogram.element_matrices(conn0, xy, conn1.shape[1], 2).cpu().numpy()
        assert rel_max(me, want) < 1e-13
        assert np.all(me[:, 0::2, 1::2] == 0.0) and np.all(me[:, 1::2, 0::2] == 0.0)
        assert rel_max(me, orc.mass_matrices(conn1 - 1, X)) < 1e-13


@pytest.mark.parametrize("kind,n,shuffle", [(4, 37, False), (4, 29, True), (9, 17, True), (3, 31, True)])
def test_mass_assembly_all_kernels(torch_cuda, kind, n, shuffle):
    """Assembled M: same pattern as K, values vs the oracle, tiled == row kernel bit for bit, bit-wise
    symmetric, and 1^T M 1 = 2 * sum_e sum_g gpWei detJ (partition of unity of the shape functions)."""
    torch = torch_cuda
    from feprogram_b200 import mesh, ops
    gen = {4: mesh.structured_quad4, 9: mesh.structured_quad9, 3: mesh.structured_tri3}[kind]
    conn, coords, bc = gen(n, n + 2)
    mx, my = (2 * n, 2 * (n + 2)) if kind == 9 else (n, n + 2)
    coords = mesh.perturb_interior(coords, mx, my, 0.15, seed=5)
    if shuffle:
        conn, coords, bc, _ = mesh.shuffle_numbering(conn, coords, bc, seed=6)
    conn0, xy = dev_mesh(torch, conn, coords)
    pat = ops.symbolic(conn0, coords.shape[0], kind, 2)
    v_rows = torch.empty(pat.nnz, dtype=torch.float64, device="cuda")
    ops.assemble_values(conn0, xy, pat, None, v_rows, 2)
    plan = ops.tile_plan(conn0, xy, pat)
    assert plan is not None
    v = torch.full((pat.nnz,), float("nan"), dtype=torch.float64, device="cuda")
    ops.assemble_values(conn0, xy, pat, plan, v, 2)
    assert torch.equal(v, v_rows)
    M = sp.csr_matrix((v.cpu().numpy(), pat.col_idx.cpu().numpy(), pat.row_ptr.cpu().numpy()), shape=(pat.nrows,) * 2)
    Mo = orc.assemble_csr(conn - 1, coords, kind, mass=True)
    assert np.array_equal(M.indptr, Mo.indptr) and np.array_equal(M.indices, Mo.indices)
    assert rel_max(M.data, Mo.data) < VAL_TOL
    assert abs(M - M.T).max() == 0.0
    dH, H, w = orc.tables(kind)
    X = coords[conn - 1]
    J = np.einsum("gka,eaj->egkj", dH, X)
    det = J[..., 0, 0] * J[..., 1, 1] - J[..., 0, 1] * J[..., 1, 0]
    total = 2.0 * float((det * w[None, :] * (H.sum(axis=1) ** 2)[None, :]).sum())
    assert abs(M.sum() - total) < 1e-11 * abs(total)


def test_mass_facade_operator(torch_cuda, golden):
    z = golden("q9_6x6_pert_shuf")
    fem = run_facade(z, operator="mass")
    Mo = orc.assemble_csr(z["conn"] - 1, z["coords"], mass=True)
    M = fem.csr_prebc()
    assert np.array_equal(M.indptr, Mo.indptr) and np.array_equal(M.indices, Mo.indices)
    assert rel_max(M.data, Mo.data) < VAL_TOL


# ---------------------------------------------------------------------------------------------------
# post-processing kernels: Elem.getHgrad at every Gauss point, nodal strain
# ---------------------------------------------------------------------------------------------------
def test_gauss_gradients_against_reference_getHgrad(torch_cuda):
    torch = torch_cuda
    import os
    from feprogram_b200.elements import FemProblem
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "post_gradients.npz"))
    for k in (4, 9):
        fem = FemProblem(g["conn%d" % k], g["coords%d" % k])
        got = fem.gaussGradients(g["U%d" % k])
        want = g["grad%d" % k]
        assert got.shape == want.shape and rel_max(got, want) < 1e-12
        eps = fem.nodalStrain(g["U%d" % k])
        assert rel_max(eps, orc.nodal_strain(g["conn%d" % k] - 1, g["coords%d" % k], g["U%d" % k])) < 1e-12


@pytest.mark.parametrize("kind,n", [(4, 300), (9, 120), (3, 257)])
def test_post_kernels_medium_and_linear_field(torch_cuda, kind, n):
    torch = torch_cuda
    from feprogram_b200 import mesh, ops
    gen = {4: mesh.structured_quad4, 9: mesh.structured_quad9, 3: mesh.structured_tri3}[kind]
    conn, coords, bc = gen(n, n + 1)
    mx, my = (2 * n, 2 * (n + 1)) if kind == 9 else (n, n + 1)
    coords = mesh.perturb_interior(coords, mx, my, 0.2, seed=8)
    conn, coords, bc, _ = mesh.shuffle_numbering(conn, coords, bc, seed=9)
    conn0, xy = dev_mesh(torch, conn, coords)
    rng = np.random.default_rng(10)
    U = rng.standard_normal(2 * coords.shape[0])
    u = torch.from_numpy(U).cuda()
    grad = torch.ops.feprogram.gauss_gradients(conn0, xy, u, kind).cpu().numpy()
    assert rel_max(grad, orc.gauss_gradients(conn - 1, coords, U, kind)) < 1e-11
    eptr, einc = torch.ops.feprogram.incidence(conn0, coords.shape[0], kind)
    eps = torch.ops.feprogram.nodal_strain(conn0, xy, u, eptr, einc, kind).cpu().numpy()
    assert rel_max(eps, orc.nodal_strain(conn - 1, coords, U, kind)) < 1e-11
    eps2 = torch.ops.feprogram.nodal_strain(conn0, xy, u, eptr, einc, kind).cpu().numpy()
    assert np.array_equal(eps, eps2)                               # fixed summation order
    A = np.array([[0.3, -1.2], [2.0, 0.7]])
    ulin = torch.from_numpy((coords @ A.T).ravel()).cuda()
    epl = torch.ops.feprogram.nodal_strain(conn0, xy, ulin, eptr, einc, kind).cpu().numpy()
    assert np.abs(epl - np.array([A[0, 0], A[1, 1], A[0, 1] + A[1, 0]])[None, :]).max() < 1e-9
    z = torch.ops.feprogram.gauss_gradients(conn0[:0], xy, u, kind)
    assert z.shape[0] == 0


# ---------------------------------------------------------------------------------------------------
# "corrected" boundary-condition mode
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind,n", [(4, 24), (9, 10)])
def test_corrected_bc_mode(torch_cuda, kind, n):
    from feprogram_b200 import mesh
    from feprogram_b200.elements import FemProblem
    gen = mesh.structured_quad4 if kind == 4 else mesh.structured_quad9
    conn, coords, bc = gen(n, n + 1)
    mx, my = (2 * n, 2 * (n + 1)) if kind == 9 else (n, n + 1)
    coords = mesh.perturb_interior(coords, mx, my, 0.15, seed=3)
    conn, coords, bc, _ = mesh.shuffle_numbering(conn, coords, bc, seed=4)
    fem = FemProblem(conn, coords)
    fem.setMatrix()
    fem.setBoundaryConditions(bc, verbose=False, mode="corrected", traction=(1.0, -0.5))
    fem.assemble()
    fem.rtol = 1e-12
    U = fem.solveProblem()
    edges0 = mesh.boundary_edges(conn, bc[3])
    assert edges0.shape == ((n + 1) if kind == 4 else (n + 1), 2 if kind == 4 else 3)
    Ks, f, u = orc.solve_corrected(conn - 1, coords, fem.dofDir, edges0, (1.0, -0.5))
    b = np.asarray(fem.brhs.todense()).ravel()
    assert np.abs(b - f).max() < 1e-14 * np.abs(f).max()
    # total load = traction * side length (left side of the unit square), minus what the Dirichlet corner took
    full = orc.edge_loads(edges0, coords, (1.0, -0.5), 2 * coords.shape[0])
    assert abs(full[0::2].sum() - 1.0) < 1e-13 and abs(full[1::2].sum() + 0.5) < 1e-13
    K = fem.K.tocsr()
    assert abs(K - K.T).max() == 0.0
    assert rel_max(K.toarray(), Ks.toarray()) < VAL_TOL
    assert np.linalg.norm(U - u) / np.linalg.norm(u) < SOL_TOL
    assert np.all(U[np.asarray(fem.dofDir)] == 0.0)


def test_corrected_mode_is_mesh_independent(torch_cuda):
    """The reference's nodal load makes |U| grow with refinement; the consistent load
    does not."""
    from feprogram_b200 import mesh
    from feprogram_b200.elements import FemProblem
    tip = []
    for n in (16, 32):
        conn, coords, bc = mesh.structured_quad4(n)
        fem = FemProblem(conn, coords)
        fem.setMatrix()
        fem.setBoundaryConditions(bc, verbose=False, mode="corrected")
        fem.assemble()
        U = fem.solveProblem().reshape(-1, 2)
        tip.append(U[(n + 1) * n])            # top-left corner node
    assert np.abs(tip[1] - tip[0]).max() < 0.05 * np.abs(tip[0]).max()


# ---------------------------------------------------------------------------------------------------
# solver-internal locality renumbering (BASELINE cfg5: shuffled node numbering)
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["q4_12x12_pert_shuf", "q9_6x6_pert_shuf"])
def test_renumbered_solver_keeps_reference_numbering(torch_cuda, golden, name):
    z = golden(name)
    a = run_facade(z, renumber=False)
    b = run_facade(z, renumber=True)
    assert b._dev["perm"] is not None and a._dev["perm"] is None
    Ka, Kb = a.csr_prebc(), b.csr_prebc()
    assert np.array_equal(Ka.indptr, Kb.indptr) and np.array_equal(Ka.indices, Kb.indices)
    assert np.array_equal(Ka.data, Kb.data)                      # same element order, same pair arithmetic
    assert np.array_equal(a.brhs.toarray(), b.brhs.toarray())
    assert abs(a.K - b.K).max() == 0.0
    for fem in (a, b):
        fem.rtol = 1e-12
    Ua, Ub = a.solveProblem(), b.solveProblem()
    assert np.linalg.norm(Ub - z["U"]) / np.linalg.norm(z["U"]) < SOL_TOL
    assert np.linalg.norm(Ub - Ua) / np.linalg.norm(Ua) < 1e-9


def test_renumbering_is_automatic_for_shuffled_large_meshes(torch_cuda):
    from feprogram_b200 import mesh
    from feprogram_b200.elements import FemProblem
    conn, coords, bc = mesh.structured_quad4(360)
    fem = FemProblem(conn, coords)
    fem.setMatrix(); fem.setBoundaryConditions(bc, verbose=False); fem.assemble()
    assert fem._dev["perm"] is None                              # row-major lattice: already local
    coords_p = mesh.perturb_interior(coords, 360, 360, 0.2, seed=0)
    conn2, coords2, bc2, _ = mesh.shuffle_numbering(conn, coords_p, bc, seed=1)
    fem2 = FemProblem(conn2, coords2)
    fem2.setMatrix(); fem2.setBoundaryConditions(bc2, verbose=False, mode="corrected"); fem2.assemble()
    assert fem2._dev["perm"] is not None
    U2 = fem2.solveProblem()
    fem3 = FemProblem(conn2, coords2, renumber=False)
    fem3.setMatrix(); fem3.setBoundaryConditions(bc2, verbose=False, mode="corrected"); fem3.assemble()
    U3 = fem3.solveProblem()
    assert fem2.solver_info["converged"] and np.linalg.norm(U2 - U3) / np.linalg.norm(U3) < 1e-7
    K2 = fem2.csr_prebc()
    Ko = orc.assemble_csr(conn2 - 1, coords2)
    assert np.array_equal(K2.indptr, Ko.indptr) and np.array_equal(K2.indices, Ko.indices)
    assert rel_max(K2.data, Ko.data) < VAL_TOL


def test_main_driver_runs_the_reference_call_order(torch_cuda, tmp_path, caplog):
    """feprogram_b200.main.run = the reference's main.py (mesh -> FemProblem -> setMatrix -> BCs -> assemble ->
    solve -> .vtu) on the CUDA path, with the reference's log labels."""
    import logging
    from feprogram_b200 import main as drv
    with caplog.at_level(logging.INFO):
        U = drv.run(str(tmp_path / "out"), nx=4, ny=6)
    assert U.shape == (35, 2) and np.isfinite(U).all() and np.all(U[:5] == [[1.0, 1.0]] + [[0.0, 0.0]] * 4)
    text = (tmp_path / "out.vtu").read_text()
    assert 'Name="Velocidad"' in text and 'NumberOfPoints="35" NumberOfCells="24"' in text
    for label in ("Crear malla", "Ensamblaje de matrices", "Calculo de velocidades", "Tiempo total"):
        assert any(label in r.getMessage() for r in caplog.records)


# ---------------------------------------------------------------------------------------------------------
# full-size parity of the BENCHMARKED kernel: the tiled kernel at the measured configurations, where every
# persistent CTA runs hundreds of tiles (steady-state double buffering), against the independent row kernel
# ---------------------------------------------------------------------------------------------------------
def _strip_oracle_check(torch, pat, vals, gen, kind, n, j0, coords_full, mx, my):
    """Rows of the nodes in lattice rows j0+1 .. j0+2 depend only on element rows j0 .. j0+2: compare them with the
    oracle run on that strip (same coordinates), structure bit-exact and values to VAL_TOL."""
    w = mx + 1
    per = 2 if kind == 9 else 1                              # lattice rows per element row
    sub_conn, sub_coords, _ = gen(n, 3)
    nsub = sub_coords.shape[0]
    base = j0 * per * w                                      # first lattice node of the strip inside the full mesh
    sub_coords = coords_full[base:base + nsub].copy()
    Ko = orc.assemble_csr(sub_conn - 1, sub_coords, kind)
    g0, g1 = 2 * ((j0 * per + per) * w), 2 * ((j0 * per + 2 * per) * w)   # global rows: lattice rows of element row j0+1
    rp = pat.row_ptr[g0:g1 + 1].cpu().numpy().astype(np.int64)
    got = vals[rp[0]:rp[-1]].cpu().numpy()
    s0, s1 = g0 - 2 * base, g1 - 2 * base
    want = Ko.data[Ko.indptr[s0]:Ko.indptr[s1]]
    assert got.shape == want.shape and rel_max(got, want) < VAL_TOL
    cols = pat.col_idx[rp[0]:rp[-1]].cpu().numpy().astype(np.int64)
    assert np.array_equal(cols - 2 * base, Ko.indices[Ko.indptr[s0]:Ko.indptr[s1]])


@pytest.mark.parametrize("kind,n,perturb", [(4, 4096, False), (4, 4096, True), (9, 1024, False), (3, 3500, True)])
def test_full_size_tiled_equals_rows_bitwise(torch_cuda, kind, n, perturb):
    """cfg2 (Quad4 4096^2, structured and perturbed), Quad9 1024^2 and Tri3 3500^2 x 2: the tiled kernel that
    bench.py times, on a NaN-prefilled output, equals the row kernel bit for bit; a strip of its rows equals
    the oracle on the sub-mesh that determines them."""
    torch = torch_cuda
    from feprogram_b200 import mesh, ops
    gen = {4: mesh.structured_quad4, 9: mesh.structured_quad9, 3: mesh.structured_tri3}[kind]
    conn, coords, _ = gen(n, dtype=np.int32)
    mx = my = 2 * n if kind == 9 else n
    if perturb:
        coords = mesh.perturb_interior(coords, mx, my, 0.2, seed=7)
    conn0 = torch.from_numpy(conn - 1).cuda()
    xy = torch.from_numpy(coords).cuda()
    del conn
    pat = ops.symbolic(conn0, coords.shape[0], kind, 2)
    plan = ops.tile_plan(conn0, xy, pat)
    assert plan is not None, "the benchmark mesh must get the tiled kernel"
    assert plan.warp_specialised == (kind == 4), "Quad4 benchmark meshes must get the warp-specialised kernel"
    v_tiled = torch.full((pat.nnz,), float("nan"), dtype=torch.float64, device="cuda")
    ops.assemble_values(conn0, xy, pat, plan, v_tiled)
    assert not bool(torch.isnan(v_tiled).any()), "an entry of K was never written"
    v_rows = torch.empty(pat.nnz, dtype=torch.float64, device="cuda")
    ops.assemble_values(conn0, xy, pat, None, v_rows)
    assert torch.equal(v_tiled, v_rows)
    del v_rows
    v2 = torch.full((pat.nnz,), float("nan"), dtype=torch.float64, device="cuda")
    ops.assemble_values(conn0, xy, pat, plan, v2)
    assert torch.equal(v_tiled, v2), "reruns must be bit-identical"
    del v2
    _strip_oracle_check(torch, pat, v_tiled, gen, kind, n, n // 3, coords, mx, my)


def test_pcg_breakdown_is_reported_not_returned_as_converged(torch_cuda):
    """An indefinite system (elements of mixed orientation: half of them have negative Jacobians, so their
    element matrices are negative semi-definite) must not come back as a converged solution."""
    torch = torch_cuda
    from feprogram_b200 import mesh
    from feprogram_b200.elements import FemProblem
    conn, coords, bc = mesh.structured_quad4(12)
    conn = conn.copy()
    conn[::2] = conn[::2, ::-1]                      # clockwise: negative determinant
    fem = FemProblem(conn, coords)
    fem.setMatrix()
    fem.setBoundaryConditions(bc, verbose=False)
    fem.assemble()
    fem.maxit = 2000
    try:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            U = fem.solveProblem()
    except RuntimeError as e:
        assert "breakdown" in str(e)
        assert fem.solver_info["breakdown"] and not fem.solver_info["converged"]
    else:       # CG may also simply fail to converge on an indefinite matrix: then it must say so
        assert not fem.solver_info["converged"] or \
            np.linalg.norm(fem.K.tocsr() @ U - np.asarray(fem.brhs.todense()).ravel()) < 1e-6


@pytest.mark.slow
def test_cfg2_pcg_runs_to_convergence(torch_cuda):
    """BASELINE config 2 solved to convergence once (about 7 n iterations): the true residual of the returned
    iterate, recomputed with fe_spmv on the unmodified matrix, meets the tolerance."""
    torch = torch_cuda
    from feprogram_b200 import mesh
    from feprogram_b200.elements import FemProblem
    n = 4096
    conn, coords, bc = mesh.structured_quad4(n, dtype=np.int32)
    fem = FemProblem(conn, coords)
    fem.setMatrix()
    fem.setBoundaryConditions(bc, verbose=False)
    fem.assemble()
    fem.rtol = 1e-9
    fem.maxit = 60000
    U = fem.solveProblem()
    info = fem.solver_info
    assert info["converged"] and 5 * n < info["iterations"] < 12 * n, info
    st = fem._dev
    pat = st["pattern"]
    x = torch.from_numpy(U).cuda()
    y = torch.empty_like(x)
    torch.ops.feprogram.spmv(pat.row_ptr, pat.col_idx, st["vals"], x, y, None)
    free = st["dirmask"] == 0
    r = (st["rhs"] - y)[free]
    b = st["rhs"].clone()
    # right-hand side of the free block: b_I - K_ID g_D
    g = torch.where(free, torch.zeros_like(x), x)
    torch.ops.feprogram.spmv(pat.row_ptr, pat.col_idx, st["vals"], g, y, None)
    b0 = (b - y)[free]
    assert float(r.norm() / b0.norm()) <= 1e-8
    assert torch.equal(x[~free], st["rhs"][~free])           # Dirichlet dofs hold their values exactly


def test_lattice_numbering_plans_without_coordinates(torch_cuda):
    """A row-major lattice numbering (structured meshes, whatever their geometry) gets its tiles from the numbering:
    tile_plan never asks for the coordinates (the facade overlaps their upload with the plan), the values are those
    of the row kernel bit for bit; any other numbering falls back to centroid keys and does ask."""
    torch = torch_cuda
    from feprogram_b200 import mesh, ops
    conn, coords, bc = mesh.structured_quad4(200, 150)
    coords = mesh.perturb_interior(coords, 200, 150, 0.25, seed=5)
    for shuffle in (False, True):
        c, x = conn, coords
        if shuffle:
            c, x, _, _ = mesh.shuffle_numbering(conn, coords, bc, seed=9)
        conn0, xy = dev_mesh(torch, c, x)
        pat = ops.symbolic(conn0, x.shape[0], 4, 2)
        asked = []
        plan = ops.tile_plan(conn0, xy, pat, coords_ready=lambda: asked.append(1))
        assert plan is not None
        assert (plan.stats["mode"] == "lattice") == (not shuffle)
        assert bool(asked) == shuffle
        v = torch.full((pat.nnz,), float("nan"), dtype=torch.float64, device="cuda")
        ops.assemble_values(conn0, xy, pat, plan, v)
        v_rows = torch.empty(pat.nnz, dtype=torch.float64, device="cuda")
        ops.assemble_values(conn0, xy, pat, None, v_rows)
        assert torch.equal(v, v_rows)
        if not shuffle:                      # exact 8 x 16 tiles: 200 x 150 elements -> 13 x 19 tiles
            assert plan.ntiles == 13 * 19


@pytest.mark.parametrize("rank,world", [(0, 2), (1, 2), (1, 3), (2, 3)])
def test_row_block_local_meshes_plan_without_coordinates(torch_cuda, rank, world):
    """The local mesh of a rank of a row-block partition numbers its owned node rows first and the ghost rows after
    them: not a row-major lattice, but every element still spans two node rows of W consecutive ids -- the generalised
    form of fe_lattice_keys.  Its tile plan needs no coordinates either (multi-GPU e2e: the plan overlaps their upload)."""
    torch = torch_cuda
    from feprogram_b200 import dist as fdist
    from feprogram_b200 import ops
    m, _ = fdist.structured_quad4_block(96, 40, rank, world)
    conn0 = torch.from_numpy(np.ascontiguousarray(m.conn)).cuda()
    xy = torch.from_numpy(np.ascontiguousarray(m.coords)).cuda()
    pat = ops.symbolic(conn0, m.nnode, 4, 2)
    asked = []
    plan = ops.tile_plan(conn0, xy, pat, coords_ready=lambda: asked.append(1))
    assert plan is not None and plan.stats["mode"] == "lattice" and not asked
    v = torch.full((pat.nnz,), float("nan"), dtype=torch.float64, device="cuda")
    ops.assemble_values(conn0, xy, pat, plan, v)
    v_rows = torch.empty(pat.nnz, dtype=torch.float64, device="cuda")
    ops.assemble_values(conn0, xy, pat, None, v_rows)
    assert torch.equal(v, v_rows)


# ---- scalar Poisson extension (FE_OP_POISSON, one dof per node; SURVEY "facts at a glance", §8d) -------------------
@pytest.mark.parametrize("name", FULL_CASES)
def test_poisson_operator_is_the_even_block_of_the_reference_matrix(torch_cuda, golden, name):
    """Pattern = node adjacency, bit-exact; values = K[0::2, 0::2] of the reference's frozen matrix (<= 1e-12) and of
    this library's own two-dof assembly (bit for bit); row kernel == tiled kernel bit for bit; Ke = Ke2[0::2, 0::2]."""
    torch = torch_cuda
    from feprogram_b200 import ops
    z = golden(name)
    conn0, xy = dev_mesh(torch, z["conn"], z["coords"])
    nnode, et = z["coords"].shape[0], conn0.shape[1]
    n2 = len(z["k0_indptr"]) - 1
    Kref = sp.csr_matrix((z["k0_data"], z["k0_indices"], z["k0_indptr"]), shape=(n2, n2))
    Ks_ref = orc.poisson_submatrix(Kref)
    pat = ops.symbolic(conn0, nnode, et, 1)
    assert pat.row_ptr.dtype == torch.int32 and pat.nnz == Ks_ref.nnz
    assert np.array_equal(pat.row_ptr.cpu().numpy(), Ks_ref.indptr)
    assert np.array_equal(pat.col_idx.cpu().numpy(), Ks_ref.indices)
    v_rows = torch.full((pat.nnz,), float("nan"), dtype=torch.float64, device="cuda")
    ops.assemble_values(conn0, xy, pat, None, v_rows, ops.FE_OP_POISSON)
    assert rel_max(v_rows.cpu().numpy(), Ks_ref.data) < VAL_TOL
    # the same bits as the L entries of the two-dof assembly
    pat2 = ops.symbolic(conn0, nnode, et, 2)
    v2 = torch.empty(pat2.nnz, dtype=torch.float64, device="cuda")
    ops.assemble_values(conn0, xy, pat2, None, v2, ops.FE_OP_REFERENCE)
    K2 = sp.csr_matrix((v2.cpu().numpy(), pat2.col_idx.cpu().numpy(), pat2.row_ptr.cpu().numpy()), shape=(n2, n2))
    assert np.array_equal(orc.poisson_submatrix(K2).data, v_rows.cpu().numpy())
    for kernel in ("auto", "phased", "ws"):          # warp-specialised (Quad4 default) and barrier-phased kernels
        plan = ops.tile_plan(conn0, xy, pat, kernel=kernel)
        if plan is None:            # no warp-specialised kernel for Quad9; irregular tiles may exceed its record slot
            assert kernel == "ws"
            continue
        assert plan.warp_specialised == (kernel == "ws") or kernel == "auto"
        v_tiled = torch.full((pat.nnz,), float("nan"), dtype=torch.float64, device="cuda")
        ops.assemble_values(conn0, xy, pat, plan, v_tiled, ops.FE_OP_POISSON)
        assert torch.equal(v_tiled, v_rows), kernel
    ke1 = torch.ops.feprogram.element_matrices(conn0[:8].contiguous(), xy, et, ops.FE_OP_POISSON).cpu().numpy()
    ke2 = torch.ops.feprogram.element_matrices(conn0[:8].contiguous(), xy, et, ops.FE_OP_REFERENCE).cpu().numpy()
    assert np.array_equal(ke1, ke2[:, 0::2, 0::2])
    with pytest.raises(ValueError):
        ops.assemble_values(conn0, xy, pat, None, v_rows, ops.FE_OP_REFERENCE)


@pytest.mark.parametrize("kind,n,perturb,shuffle", [(4, 96, True, False), (4, 200, True, True), (9, 40, True, False),
                                                     (3, 120, True, True)])
def test_poisson_tiled_equals_rows_on_larger_meshes(torch_cuda, kind, n, perturb, shuffle):
    torch = torch_cuda
    from feprogram_b200 import mesh, ops
    gen = {4: mesh.structured_quad4, 9: mesh.structured_quad9, 3: mesh.structured_tri3}[kind]
    conn, coords, bc = gen(n)
    mx = 2 * n if kind == 9 else n
    if perturb:
        coords = mesh.perturb_interior(coords, mx, mx, 0.2, seed=3)
    if shuffle:
        conn, coords, bc, _ = mesh.shuffle_numbering(conn, coords, bc, seed=4)
    conn0, xy = dev_mesh(torch, conn, coords)
    pat = ops.symbolic(conn0, coords.shape[0], kind, 1)
    v_rows = torch.full((pat.nnz,), float("nan"), dtype=torch.float64, device="cuda")
    ops.assemble_values(conn0, xy, pat, None, v_rows, ops.FE_OP_POISSON)
    for kernel in ("auto", "phased") + (() if kind == 9 else ("ws",)):
        plan = ops.tile_plan(conn0, xy, pat, kernel=kernel)
        assert plan is not None, kernel
        G = 4096                                             # guard zones either side of the output stay untouched
        buf = torch.full((pat.nnz + 2 * G,), float("nan"), dtype=torch.float64, device="cuda")
        v = buf[G:G + pat.nnz]
        ops.assemble_values(conn0, xy, pat, plan, v, ops.FE_OP_POISSON)
        assert torch.equal(v, v_rows), kernel
        assert bool(torch.isnan(buf[:G]).all()) and bool(torch.isnan(buf[G + pat.nnz:]).all()), kernel
    # constants are in the null space of the scalar stiffness; the matrix is symmetric bit for bit
    K = sp.csr_matrix((v.cpu().numpy(), pat.col_idx.cpu().numpy(), pat.row_ptr.cpu().numpy()),
                      shape=(pat.nrows, pat.nrows))
    assert np.abs(K @ np.ones(pat.nrows)).max() < 1e-11 * np.abs(K.data).max()
    assert abs(K - K.T).max() == 0.0
    Ks = orc.poisson_submatrix(orc.assemble_csr(conn - 1, coords, kind))
    assert np.array_equal(K.indptr, Ks.indptr) and np.array_equal(K.indices, Ks.indices)
    assert rel_max(K.data, Ks.data) < VAL_TOL


@pytest.mark.parametrize("name", FULL_CASES)
def test_poisson_facade_solution(torch_cuda, golden, name):
    """FemProblem(operator="poisson"): same call order, one dof per node, the reference's BC semantics on nodes."""
    z = golden(name)
    fem = run_facade(z, operator="poisson")
    nn = fem.nnode
    Ks, Kbc, b, u = orc.solve_poisson_path(z["conn"] - 1, z["coords"], bc_sets_of(z))
    K0 = fem.csr_prebc()
    assert K0.shape == (nn, nn) and np.array_equal(K0.indptr, Ks.indptr) and np.array_equal(K0.indices, Ks.indices)
    assert rel_max(K0.data, Ks.data) < VAL_TOL
    assert np.array_equal(np.asarray(fem.brhs.todense()).ravel(), b)
    assert rel_max(fem.K.toarray(), Kbc.toarray()) < VAL_TOL
    U = fem.solveProblem()
    assert U.shape == (nn,) and fem.solver_info["converged"]
    assert np.linalg.norm(U - u) / np.linalg.norm(u) < SOL_TOL
    with pytest.raises(ValueError):
        fem.setBoundaryConditions(bc_sets_of(z), verbose=False, mode="corrected")


# ---- genuinely unstructured meshes (varying node degree, random numbering): what gmsh hands the reference ----------
@pytest.mark.parametrize("kind,n", [(4, 60), (3, 150)])
def test_unstructured_meshes_structure_values_kernels_and_solution(torch_cuda, kind, n):
    torch = torch_cuda
    from feprogram_b200 import mesh, ops
    from feprogram_b200.elements import FemProblem
    gen = mesh.unstructured_quad4 if kind == 4 else mesh.unstructured_tri3
    conn, coords, bc = gen(n, seed=9)
    conn0, xy = dev_mesh(torch, conn, coords)
    nnode = coords.shape[0]
    Ko = orc.assemble_csr(conn - 1, coords, kind)
    pat = ops.symbolic(conn0, nnode, kind, 2)
    assert np.array_equal(pat.row_ptr.cpu().numpy(), Ko.indptr) and np.array_equal(pat.col_idx.cpu().numpy(), Ko.indices)
    assert pat.max_deg > 9                               # the degree really varies
    v_rows = torch.full((pat.nnz,), float("nan"), dtype=torch.float64, device="cuda")
    ops.assemble_values(conn0, xy, pat, None, v_rows)
    assert rel_max(v_rows.cpu().numpy(), Ko.data) < VAL_TOL
    used = []
    for kernel in ("auto", "phased", "ws"):
        plan = ops.tile_plan(conn0, xy, pat, kernel=kernel)
        if plan is None:                                 # a capacity does not hold: the facade then uses the row kernel
            continue
        used.append((kernel, plan.stats.get("mode"), plan.warp_specialised))
        v = torch.full((pat.nnz,), float("nan"), dtype=torch.float64, device="cuda")
        ops.assemble_values(conn0, xy, pat, plan, v)
        assert torch.equal(v, v_rows), (kernel, plan.stats)
    assert any(k == "phased" for k, _, _ in used), used   # the tiled path must exist for irregular meshes
    if kind == 4:                                        # whole path through the facade (the reference has no Tri3)
        fem = FemProblem(conn, coords)
        fem.setMatrix()
        fem.setBoundaryConditions(bc, verbose=False)
        fem.assemble()
        U = fem.solveProblem()
        _, _, b, Uo = orc.solve_reference_path(conn - 1, coords, bc)
        assert np.array_equal(fem.rhs_host(), b)
        assert fem.solver_info["converged"]
        assert np.linalg.norm(U - Uo) / np.linalg.norm(Uo) < SOL_TOL


# ---- nodes of very high valence: every capacity of the symbolic write-out (32 staged neighbours, 16 staged incidences)
# is exceeded by a hub, the nodes around it stay on the fast path -- both in one warp's flat write-out -------------------
def _fan_mesh(spokes, hubs=3, seed=3):
    """Tri3 fans: `hubs` hub nodes, each surrounded by a closed ring of `spokes` triangles; a strip of regular
    triangles joins consecutive rings so that the mesh is connected.  Node ids are shuffled."""
    rng = np.random.default_rng(seed)
    coords, tris = [], []
    for h in range(hubs):
        base = len(coords)
        cx = 3.0 * h
        coords.append((cx, 0.0))
        ang = 2 * np.pi * np.arange(spokes) / spokes
        coords.extend(zip(cx + np.cos(ang), np.sin(ang)))
        for i in range(spokes):
            tris.append((base, base + 1 + i, base + 1 + (i + 1) % spokes))
        if h > 0:                                   # a triangle pair between the previous ring and this one
            prev = base - (spokes + 1)
            tris.append((prev + 1, base + 1 + spokes // 2, base + 1 + spokes // 2 + 1))
            tris.append((prev + 1, base + 1 + spokes // 2 + 1, prev + 2))
    coords = np.asarray(coords, dtype=np.float64)
    perm = rng.permutation(len(coords))
    inv = np.empty_like(perm)
    inv[perm] = np.arange(len(coords))
    conn = inv[np.asarray(tris, dtype=np.int64)] + 1
    return conn[rng.permutation(len(conn))].astype(np.int32), coords[perm]


@pytest.mark.parametrize("spokes", [12, 20, 40, 70])
def test_high_valence_hubs_symbolic_and_values(torch_cuda, spokes):
    torch = torch_cuda
    from feprogram_b200 import ops
    conn, coords = _fan_mesh(spokes)
    conn0, xy = dev_mesh(torch, conn, coords)
    nnode = coords.shape[0]
    pat = ops.symbolic(conn0, nnode, 3, 2, csr="lazy")
    assert pat._col_idx is None                      # nothing expanded yet
    eptr, einc = orc.node_incidence(conn - 1, nnode)
    nptr, nadj = orc.node_adjacency(conn - 1, nnode)
    assert np.array_equal(pat.eptr.cpu().numpy(), eptr) and np.array_equal(pat.einc.cpu().numpy(), einc)
    assert np.array_equal(pat.nptr.cpu().numpy(), nptr) and np.array_equal(pat.nadj.cpu().numpy(), nadj)
    assert pat.max_deg == spokes + 1
    epos = pat.epos.cpu().numpy().reshape(-1, 3)     # every incidence: positions of the element's nodes in the owner's list
    owner = np.repeat(np.arange(nnode), np.diff(eptr))
    for k in range(len(einc)):
        seg = nadj[nptr[owner[k]]:nptr[owner[k] + 1]]
        assert np.array_equal(seg[epos[k]], (conn - 1)[einc[k] // 16])
    Ko = orc.assemble_csr(conn - 1, coords, 3)
    assert np.array_equal(pat.row_ptr.cpu().numpy(), Ko.indptr) and np.array_equal(pat.col_idx.cpu().numpy(), Ko.indices)
    v = torch.full((pat.nnz,), float("nan"), dtype=torch.float64, device="cuda")
    ops.assemble_values(conn0, xy, pat, None, v)     # hubs exceed the tile plans' capacities: row kernel
    assert rel_max(v.cpu().numpy(), Ko.data) < VAL_TOL


def test_lazy_and_eager_csr_patterns_are_the_same(torch_cuda, golden):
    torch = torch_cuda
    from feprogram_b200 import ops
    z = golden(FULL_CASES[0])
    conn0, _ = dev_mesh(torch, z["conn"], z["coords"])
    nnode = z["coords"].shape[0]
    eager = ops.symbolic(conn0, nnode, z["conn"].shape[1], 2)
    lazy = ops.symbolic(conn0, nnode, z["conn"].shape[1], 2, csr="lazy")
    assert lazy._row_ptr is None and lazy.nnz == eager.nnz == len(z["k0_indices"])
    assert torch.equal(lazy.row_ptr, eager.row_ptr) and torch.equal(lazy.col_idx, eager.col_idx)
    assert np.array_equal(lazy.col_idx.cpu().numpy(), z["k0_indices"])
