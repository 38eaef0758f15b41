"""Host-side logic that needs no GPU: numbering locality, boundary edges, the corrected-load oracle, the facade's
argument checking.  (The kernels behind them are covered by the -m gpu tests.)"""
import numpy as np
import pytest
import torch

from feprogram_b200 import mesh
from oracle import fe_oracle as orc


def test_locality_order_is_a_permutation_with_spatial_locality():
    from feprogram_b200 import ops
    conn, coords, _ = mesh.structured_quad4(64)
    coords = mesh.perturb_interior(coords, 64, 64, 0.2, seed=1)
    conn2, coords2, _, _ = mesh.shuffle_numbering(conn, coords, [set()] * 4, seed=2)
    xy = torch.from_numpy(coords2)
    perm, order = ops.locality_order(xy)
    n = coords2.shape[0]
    assert torch.equal(torch.sort(perm).values, torch.arange(n)) and torch.equal(perm[order], torch.arange(n))
    c0 = torch.from_numpy((conn2 - 1).astype(np.int32))
    span_in = ops.numbering_span(c0)
    span_out = ops.numbering_span(perm[c0.long()].to(torch.int32))
    assert span_in > 64.0 * np.sqrt(n) / 8 and span_out < span_in / 10          # shuffled -> local
    # a row-major lattice is left alone by the automatic rule (span ~ row length)
    assert ops.numbering_span(torch.from_numpy((conn - 1).astype(np.int32))) <= 66 + 1


def test_partition_morton_keys_match_device_convention():
    from feprogram_b200 import dist as fdist
    xy = np.random.default_rng(0).random((500, 2))
    k = fdist.morton_keys(xy, xy.min(axis=0), xy.max(axis=0) - xy.min(axis=0))
    assert k.min() >= 0 and k.max() < 2 ** 32 and len(np.unique(k)) > 490
    # points in the same quadrant share the two top key bits
    q = ((xy[:, 0] > 0.5 * (xy[:, 0].min() + xy[:, 0].max())).astype(int)
         + 2 * (xy[:, 1] > 0.5 * (xy[:, 1].min() + xy[:, 1].max())).astype(int))
    assert np.array_equal(k >> 30, q)


@pytest.mark.parametrize("kind", [4, 9, 3])
def test_boundary_edges_of_each_side(kind):
    gen = {4: mesh.structured_quad4, 9: mesh.structured_quad9, 3: mesh.structured_tri3}[kind]
    nx, ny = 5, 3
    conn, coords, bc = gen(nx, ny)
    conn, coords, bc, _ = mesh.shuffle_numbering(conn, coords, bc, seed=3)
    want = [nx, ny, nx, ny]                       # bottom, right, top, left: one edge per element side
    for side, cnt in zip(bc, want):
        e = mesh.boundary_edges(conn, side)
        assert e.shape == (cnt, 3 if kind == 9 else 2) and e.dtype == np.int32
        tags = np.fromiter((int(t) for t in side), dtype=np.int64) - 1
        assert np.isin(e, tags).all()
        # consistent load of a unit traction sums to the side length (1.0 on the unit square)
        f = orc.edge_loads(e, coords, (1.0, 2.0), 2 * coords.shape[0])
        assert abs(f[0::2].sum() - 1.0) < 1e-13 and abs(f[1::2].sum() - 2.0) < 1e-13
    assert mesh.boundary_edges(conn, []).shape[0] == 0


def test_quadratic_edge_load_weights():
    # straight 3-node edge of length 2: (1/6, 1/6, 4/6) * length
    coords = np.array([[0.0, 0.0], [2.0, 0.0], [1.0, 0.0]])
    f = orc.edge_loads(np.array([[0, 1, 2]]), coords, (1.0, 0.0), 6)
    assert np.allclose(f[0::2], [2 / 6, 2 / 6, 8 / 6], atol=1e-14)


def test_facade_rejects_bad_arguments_without_touching_the_gpu():
    from feprogram_b200.elements import FemProblem
    conn, coords, bc = mesh.structured_quad4(3)
    with pytest.raises(KeyError):
        FemProblem(conn, coords, operator="nonsense")
    fem = FemProblem(conn, coords)
    fem.setMatrix()
    with pytest.raises(ValueError):
        fem.setBoundaryConditions(bc, verbose=False, mode="symmetric")
    fem.setBoundaryConditions(bc, verbose=False, mode="corrected")
    assert not set(fem.dofNeu) & set(fem.dofDir)            # Dirichlet wins at the shared corner
    fem.setBoundaryConditions(bc, verbose=False)
    assert set(fem.dofNeu) & set(fem.dofDir)                # the reference's overlap is kept in reference mode
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            fem.assemble()


@pytest.mark.parametrize("seed", range(6))
def test_partition_invariants_random_meshes(seed):
    """partition_mesh on shuffled meshes of every element type, 2..5 ranks, both local numberings: every node
    owned exactly once, every rank holds all elements touching its owned nodes, local ids map back to the
    global mesh, ghosts are grouped by owner, element blocks tile the element range."""
    from feprogram_b200 import dist as fdist
    rng = np.random.default_rng(seed)
    kind = (4, 9, 3)[seed % 3]
    gen = {4: mesh.structured_quad4, 9: mesh.structured_quad9, 3: mesh.structured_tri3}[kind]
    nx, ny = int(rng.integers(3, 9)), int(rng.integers(3, 9))
    conn, coords, bc = gen(nx, ny)
    conn, coords, bc, _ = mesh.shuffle_numbering(conn, coords, bc, seed=seed + 10)
    conn0 = conn - 1
    world = int(rng.integers(2, 6))
    locality = bool(seed & 1)
    bounds = fdist.element_blocks(conn0.shape[0], world)
    assert bounds[0] == 0 and bounds[-1] == conn0.shape[0] and np.all(np.diff(bounds) >= 0)
    owned_count = np.zeros(coords.shape[0], dtype=int)
    for rank in range(world):
        m = fdist.partition_mesh(conn0, coords, rank, world, locality=locality)
        owned = m.node_gid[: m.n_owned]
        owned_count[owned] += 1
        # local connectivity maps back to the global one, elements in ascending global id, own block first-class
        assert np.array_equal(m.node_gid[m.conn], conn0[m.elem_gid])
        assert np.all(np.diff(m.elem_gid) > 0)
        assert m.n_own_elems == int(((m.elem_gid >= bounds[rank]) & (m.elem_gid < bounds[rank + 1])).sum())
        assert np.array_equal(m.coords, coords[m.node_gid, :2])
        # every element touching an owned node is present
        touching = np.nonzero(np.isin(conn0, owned).any(axis=1))[0]
        assert np.array_equal(touching, m.elem_gid)
        # owner of a node = rank of the block holding its lowest incident element; ghosts grouped by owner
        first = np.full(coords.shape[0], conn0.shape[0])
        np.minimum.at(first, conn0.ravel(), np.repeat(np.arange(conn0.shape[0]), conn0.shape[1]))
        own_rank = np.searchsorted(bounds, first, side="right") - 1
        assert np.all(own_rank[owned] == rank)
        ghosts = m.node_gid[m.n_owned:]
        assert np.array_equal(own_rank[ghosts], m.ghost_owner) and np.all(np.diff(m.ghost_owner) >= 0)
        assert not np.any(m.ghost_owner == rank)
        if not locality:
            assert np.all(np.diff(owned) > 0)
    assert np.all(owned_count == 1)


# ---- the hand-written sorting networks of the symbolic kernels (csrc/fe_symbolic.cu), checked from their source ------
def _zero_one_sorts(pairs, n):
    """0-1 principle: a comparator network sorts every input iff it sorts every 0/1 input."""
    import itertools
    for bits in itertools.product((0, 1), repeat=n):
        v = list(bits)
        for a, b in pairs:
            if v[a] > v[b]:
                v[a], v[b] = v[b], v[a]
        if v != sorted(v):
            return False
    return True


def test_sorting_networks_in_the_symbolic_kernels_sort():
    import os
    import re
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "feprogram_b200", "csrc",
                            "fe_symbolic.cu")).read()
    # 16 packed candidate keys of a regular Quad4 node (quad4_sorted_candidates): CX(i, j) list
    body = src[src.index("#define CX(i, j)"):src.index("#undef CX")]
    pairs16 = [(int(a), int(b)) for a, b in re.findall(r"CX\((\d+), (\d+)\);", body)]
    assert len(pairs16) == 63 and all(0 <= a < b < 16 for a, b in pairs16)
    assert _zero_one_sorts(pairs16, 16)
    # eight incidence entries of a node (k_sort_incidence): cx(v[i], v[j]) list
    body = src[src.index("19-comparator network"):src.index("einc[lo + i] = v[i];")]
    pairs8 = [(int(a), int(b)) for a, b in re.findall(r"cx\(v\[(\d+)\], v\[(\d+)\]\);", body)]
    assert len(pairs8) == 19 and all(0 <= a < b < 8 for a, b in pairs8)
    assert _zero_one_sorts(pairs8, 8)


def test_warp_bitonic_network_sorts_256_keys():
    """Emulation of warp_bitonic_sort<8> (csrc/fe_tile_config.cuh): 8 keys per lane, key index = 32 r + lane,
    register exchanges at distance >= 32, shuffles below -- same index logic, random and padded inputs."""
    rng = np.random.default_rng(5)
    R = 8
    for nn in (1, 31, 32, 33, 100, 153, 255, 256):
        ids = rng.permutation(10 ** 6)[:nn]
        v = np.full((R, 32), 0x7FFFFFFF, dtype=np.int64)
        v.reshape(-1)[:nn] = ids
        lanes = np.arange(32)
        k = 2
        while k <= 32 * R:
            j = k >> 1
            while j > 0:
                if j >= 32:
                    for r in range(R):
                        q = r ^ (j >> 5)
                        if q > r:
                            up = ((32 * r) & k) == 0
                            mn, mx = np.minimum(v[r], v[q]), np.maximum(v[r], v[q])
                            v[r], v[q] = (mn, mx) if up else (mx, mn)
                else:
                    for r in range(R):
                        o = v[r][lanes ^ j]
                        up, lower = ((32 * r + lanes) & k) == 0, (lanes & j) == 0
                        v[r] = np.where(up == lower, np.minimum(v[r], o), np.maximum(v[r], o))
                j >>= 1
            k <<= 1
        assert np.array_equal(v.reshape(-1)[:nn], np.sort(ids))
