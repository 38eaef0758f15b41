"""Synthetic mesh generators for the FEM hot path.

The reference obtains ``(conectivity, coordinates, bcNodesList)`` from
``gmsh_api.getMeshInfo()`` (/root/reference/gmsh_api.py:13-113).  gmsh is not
available offline, so these generators produce the same triple:

* ``conectivity``  -- (nelem, nnpe) integer ndarray of **1-based** node tags
* ``coordinates``  -- (nnode, 2) float64 ndarray, row ``i`` <-> tag ``i + 1``
* ``bcNodesList``  -- four sets of 1-based tags ordered [bottom, right, top, left]
                      (gmsh_api.py:91-98 consumed by elements.py:207-214)

Everything here is host-side NumPy; it is input preparation, not the hot path.
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "structured_quad4",
    "structured_quad9",
    "structured_tri3",
    "perturb_interior",
    "shuffle_numbering",
    "edge_sets",
    "unstructured_tri3",
    "unstructured_quad4",
]


def _lattice(mx: int, my: int):
    """Row-major (mx+1)x(my+1) node lattice on the unit square."""
    xs = np.linspace(0.0, 1.0, mx + 1)
    ys = np.linspace(0.0, 1.0, my + 1)
    xx, yy = np.meshgrid(xs, ys)  # yy varies along axis 0 -> id = j*(mx+1)+i
    return np.ascontiguousarray(np.stack([xx.ravel(), yy.ravel()], axis=1))


def edge_sets(mx: int, my: int):
    """Boundary node tag sets of an (mx+1)x(my+1) lattice, 1-based, corners in two sets."""
    w = mx + 1
    i = np.arange(mx + 1, dtype=np.int64)
    j = np.arange(my + 1, dtype=np.int64)
    bottom = i + 1
    right = j * w + mx + 1
    top = my * w + i + 1
    left = j * w + 1
    return [set(bottom.tolist()), set(right.tolist()), set(top.tolist()), set(left.tolist())]


def structured_quad4(nx: int, ny: int | None = None, dtype=np.int64):
    """nx x ny bilinear quads, counter-clockwise tags ``[n0, n0+1, n0+nx+2, n0+nx+1] + 1``."""
    ny = nx if ny is None else ny
    coords = _lattice(nx, ny)
    i = np.arange(nx, dtype=np.int64)[None, :]
    j = np.arange(ny, dtype=np.int64)[:, None]
    n0 = (j * (nx + 1) + i).ravel()
    conn = np.stack([n0, n0 + 1, n0 + nx + 2, n0 + nx + 1], axis=1) + 1
    return conn.astype(dtype), coords, edge_sets(nx, ny)


def structured_quad9(nx: int, ny: int | None = None, dtype=np.int64):
    """nx x ny biquadratic quads on a (2nx+1)x(2ny+1) lattice.

    Local node order is the reference's (tools/interpFunctions.py:55-84):
    1(+,+) 2(-,+) 3(-,-) 4(+,-) 5(0,+) 6(-,0) 7(0,-) 8(+,0) 9(0,0).
    """
    ny = nx if ny is None else ny
    mx, my = 2 * nx, 2 * ny
    coords = _lattice(mx, my)
    w = mx + 1
    i0 = (2 * np.arange(nx, dtype=np.int64))[None, :]
    j0 = (2 * np.arange(ny, dtype=np.int64))[:, None]

    def nid(di, dj):
        return ((j0 + dj) * w + (i0 + di)).ravel()

    conn = np.stack(
        [nid(2, 2), nid(0, 2), nid(0, 0), nid(2, 0), nid(1, 2), nid(0, 1), nid(1, 0), nid(2, 1), nid(1, 1)],
        axis=1,
    ) + 1
    return conn.astype(dtype), coords, edge_sets(mx, my)


def structured_tri3(nx: int, ny: int | None = None, dtype=np.int64):
    """Each quad of the nx x ny lattice split along the (n0, n0+nx+2) diagonal, both CCW.

    Tri3 is an extension: the reference only knows Quad4/Quad9 (elements.py:55).
    """
    ny = nx if ny is None else ny
    coords = _lattice(nx, ny)
    i = np.arange(nx, dtype=np.int64)[None, :]
    j = np.arange(ny, dtype=np.int64)[:, None]
    n0 = (j * (nx + 1) + i).ravel()
    lower = np.stack([n0, n0 + 1, n0 + nx + 2], axis=1)
    upper = np.stack([n0, n0 + nx + 2, n0 + nx + 1], axis=1)
    conn = np.empty((2 * n0.size, 3), dtype=np.int64)
    conn[0::2] = lower
    conn[1::2] = upper
    return (conn + 1).astype(dtype), coords, edge_sets(nx, ny)


def perturb_interior(coords: np.ndarray, mx: int, my: int, frac: float = 0.2, seed: int = 0):
    """Interior lattice nodes += U(-frac*h, frac*h) per coordinate."""
    rng = np.random.default_rng(seed)
    out = coords.copy()
    w = mx + 1
    ids = np.arange(out.shape[0])
    i, j = ids % w, ids // w
    interior = (i > 0) & (i < mx) & (j > 0) & (j < my)
    h = np.array([1.0 / mx, 1.0 / my])
    noise = rng.uniform(-frac, frac, size=out.shape) * h
    out[interior] += noise[interior]
    return out


def shuffle_numbering(conn: np.ndarray, coords: np.ndarray, bc_sets, seed: int = 1):
    """Random renumbering of node tags; coordinates rows permuted consistently.

    ``perm[old0] = new0``: old 0-based node ``old0`` becomes node ``perm[old0]``.
    """
    rng = np.random.default_rng(seed)
    nnode = coords.shape[0]
    perm = rng.permutation(nnode)
    new_coords = np.empty_like(coords)
    new_coords[perm] = coords
    new_conn = (perm[np.asarray(conn, dtype=np.int64) - 1] + 1).astype(conn.dtype)
    new_sets = [set((perm[np.fromiter(s, dtype=np.int64, count=len(s)) - 1] + 1).tolist()) for s in bc_sets]
    return new_conn, new_coords, new_sets, perm


def boundary_edges(conn1: np.ndarray, nodes) -> np.ndarray:
    """0-based boundary edges -- element sides whose corner pair occurs in one element only -- all of whose
    nodes are in ``nodes`` (1-based tags): rows (a, b) for Quad4 / Tri3, (a, b, mid) for Quad9."""
    conn = np.asarray(conn1 if isinstance(conn1, np.ndarray) else np.stack([np.asarray(c) for c in conn1])).astype(np.int64)
    nnpe = conn.shape[1]
    nc = 3 if nnpe == 3 else 4
    a = np.concatenate([conn[:, k] for k in range(nc)])
    b = np.concatenate([conn[:, (k + 1) % nc] for k in range(nc)])
    cols = [a, b]
    if nnpe == 9:
        cols.append(np.concatenate([conn[:, nc + k] for k in range(nc)]))
    key = np.minimum(a, b) * (int(conn.max()) + 1) + np.maximum(a, b)
    _, first, count = np.unique(key, return_index=True, return_counts=True)
    bnd = np.sort(first[count == 1])
    tags = np.fromiter((int(t) for t in nodes), dtype=np.int64)
    on = np.ones(bnd.size, dtype=bool)
    for c in cols:
        on &= np.isin(c[bnd], tags)
    return (np.stack([c[bnd[on]] for c in cols], axis=1) - 1).astype(np.int32)


# ----------------------------------------------------------------------------------------
# genuinely unstructured meshes (varying node degree), the mesh class gmsh hands the reference
# ----------------------------------------------------------------------------------------
def _side_sets(coords: np.ndarray):
    """[bottom, right, top, left] 1-based tag sets of a mesh of the unit square (corners in two sets)."""
    x, y = coords[:, 0], coords[:, 1]
    pick = lambda m: set((np.nonzero(m)[0] + 1).tolist())  # noqa: E731
    return [pick(y == 0.0), pick(x == 1.0), pick(y == 1.0), pick(x == 0.0)]


def unstructured_tri3(n: int, seed: int = 0, dtype=np.int64):
    """Delaunay triangulation of the unit square: an (n+1) x (n+1) jittered point set (boundary points stay on the
    boundary, interior points move by up to 0.35 h) in RANDOM node order -- node degrees 4..9, no numbering locality.
    Counter-clockwise triangles.  Needs SciPy (host-side input preparation only)."""
    from scipy.spatial import Delaunay
    rng = np.random.default_rng(seed)
    pts = perturb_interior(_lattice(n, n), n, n, 0.35, seed=seed + 1)
    pts = pts[rng.permutation(pts.shape[0])]
    tri = Delaunay(pts).simplices.astype(np.int64)
    a, b, c = pts[tri[:, 0]], pts[tri[:, 1]], pts[tri[:, 2]]
    area2 = (b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (b[:, 1] - a[:, 1]) * (c[:, 0] - a[:, 0])
    tri = tri[np.abs(area2) > 1e-14]                       # collinear boundary slivers
    area2 = area2[np.abs(area2) > 1e-14]
    flip = area2 < 0
    tri[flip] = tri[flip][:, [0, 2, 1]]
    return (tri + 1).astype(dtype), np.ascontiguousarray(pts), _side_sets(pts)


def unstructured_quad4(n: int, seed: int = 0, dtype=np.int64):
    """All-quad unstructured mesh: every triangle of ``unstructured_tri3`` is split into three quads through its
    edge midpoints and centroid (what a recombining mesher produces locally): node degrees 3..~10, convex
    counter-clockwise quads, random node order."""
    tri, pts, _ = unstructured_tri3(n, seed, np.int64)
    tri = tri - 1
    npt = pts.shape[0]
    e = np.sort(np.concatenate([tri[:, [0, 1]], tri[:, [1, 2]], tri[:, [2, 0]]]), axis=1)
    key = e[:, 0] * npt + e[:, 1]
    uniq, inv = np.unique(key, return_inverse=True)
    mid = 0.5 * (pts[uniq // npt] + pts[uniq % npt])
    nt = tri.shape[0]
    m01, m12, m20 = (npt + inv[:nt]), (npt + inv[nt:2 * nt]), (npt + inv[2 * nt:])
    g = npt + uniq.size + np.arange(nt)
    cen = pts[tri].mean(axis=1)
    coords = np.concatenate([pts, mid, cen])
    quads = np.concatenate([np.stack([tri[:, 0], m01, g, m20], axis=1),
                            np.stack([tri[:, 1], m12, g, m01], axis=1),
                            np.stack([tri[:, 2], m20, g, m12], axis=1)])
    rng = np.random.default_rng(seed + 2)
    quads = quads[rng.permutation(quads.shape[0])]
    perm = rng.permutation(coords.shape[0])                 # old node -> new node
    new_coords = np.empty_like(coords)
    new_coords[perm] = coords
    return (perm[quads] + 1).astype(dtype), np.ascontiguousarray(new_coords), _side_sets(new_coords)
