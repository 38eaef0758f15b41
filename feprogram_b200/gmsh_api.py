"""``getMeshInfo()`` with the reference's contract (/root/reference/gmsh_api.py:13-113; next-row f-2 of
SURVEY.md §8, not on the hot path): returns ``(conectivity, coordinates, bcNodesList)`` of the 2 x 10
rectangle meshed with quads of size lc = 0.25 -- 1-based tags, coordinates (nnode, 2) (gmsh_api.py:104-106),
four boundary tag sets ordered [bottom, right, top, left].

The reference drives the gmsh SDK (an external C++ mesher).  When ``gmsh`` is importable the same model is
built with it (same calls, including the ``tp-EFAF.msh`` side-effect file); otherwise (this image has no gmsh)
a WARNING is logged and the transfinite-equivalent structured 8 x 40 quad mesh of the same rectangle is
generated, which is what gmsh's recombined mesh of this geometry amounts to.
"""
import logging

import numpy as np

from . import mesh

__all__ = ["getMeshInfo", "readMsh", "writeMsh"]

# MSH element type -> (dimension, nodes per element); 2-D node orders coincide with the reference's tables
# (corners, then mid-sides 1-2, 2-3, 3-4, 4-1, then the centre node; tools/interpFunctions.py:17-37)
_MSH_TYPES = {1: (1, 2), 8: (1, 3), 2: (2, 3), 3: (2, 4), 10: (2, 9), 15: (0, 1)}

LX, LY, LC = 2.0, 10.0, 0.25


def _structured():
    nx, ny = int(round(LX / LC)), int(round(LY / LC))
    conn, xy, bc = mesh.structured_quad4(nx, ny, dtype=np.uint64)
    coords = np.zeros((xy.shape[0], 2))
    coords[:, 0] = xy[:, 0] * LX
    coords[:, 1] = xy[:, 1] * LY
    return list(conn), coords, bc


def getMeshInfo():
    try:
        import gmsh  # noqa: F401
    except Exception:
        logging.getLogger(__name__).warning(
            "gmsh is not importable: getMeshInfo() returns the structured 8 x 40 Quad4 mesh of the 2 x 10 rectangle "
            "instead of gmsh's recombined mesh (same contract, not the same node numbering)")
        return _structured()
    gmsh.initialize()
    try:
        gmsh.option.setNumber("General.Terminal", 0)
        gmsh.model.add("t1")
        for tag, (x, y) in enumerate(((0, 0), (LX, 0), (LX, LY), (0, LY)), start=1):
            gmsh.model.geo.addPoint(x, y, 0, LC, tag)
        gmsh.model.geo.addLine(1, 2, 1)
        gmsh.model.geo.addLine(3, 2, 2)
        gmsh.model.geo.addLine(3, 4, 3)
        gmsh.model.geo.addLine(4, 1, 4)
        gmsh.model.geo.addCurveLoop([4, 1, -2, 3], 1)
        gmsh.model.geo.addPlaneSurface([1], 1)
        gmsh.model.addPhysicalGroup(0, [1, 2], 1)
        for line, (group, name) in enumerate(((2, "Bottom"), (3, "Right"), (4, "Top"), (5, "Left")), start=1):
            gmsh.model.addPhysicalGroup(1, [line], group)
            gmsh.model.setPhysicalName(1, group, name)
        gmsh.model.addPhysicalGroup(2, [1], 6)
        gmsh.model.setPhysicalName(2, 6, "Interior")
        gmsh.model.geo.mesh.setRecombine(2, 1)          # as the reference (gmsh_api.py:84): recombine surface 1
        gmsh.model.geo.synchronize()
        gmsh.model.mesh.generate(2)
        gmsh.write("tp-EFAF.msh")                       # the reference's side-effect file (gmsh_api.py:90)
        sides = []
        for line in (1, 2, 3, 4):                       # bottom, right, top, left: nodes of the 1-D elements
            _, _, enodes1 = gmsh.model.mesh.getElements(1, line)
            sides.append(set(enodes1[0]))
        etypes, _, enodes = gmsh.model.mesh.getElements(2, 1)
        nnpe = _MSH_TYPES[int(etypes[0])][1]
        conn = [np.asarray(r) for r in np.asarray(enodes[0], dtype=np.uint64).reshape(-1, nnpe)]
        tags, xyz, _ = gmsh.model.mesh.getNodes()
        coords = np.zeros((len(tags), 2))               # row order = gmsh's node order, as gmsh_api.py:104-106
        coords[:, 0] = xyz[0::3]
        coords[:, 1] = xyz[1::3]
        return conn, coords, sides
    finally:
        gmsh.finalize()


def readMsh(path, sides=(1, 2, 3, 4), surface=None):
    """Reads an ASCII MSH 4.1 file -- what the reference saves as ``tp-EFAF.msh`` (gmsh_api.py:89) -- and
    returns ``(conectivity, coordinates, bcNodesList)`` as ``getMeshInfo`` does (gmsh_api.py:90-110):
    conectivity = list of uint64 tag arrays of the 2-D elements (of ``surface``, default all surfaces),
    coordinates (max tag, 2) with row i <-> tag i + 1, bcNodesList[k] = set of the node tags of the 1-D
    elements classified on curve ``sides[k]`` ([bottom, right, top, left] in the reference's model)."""
    with open(path) as f:
        text = f.read()

    def section(name):
        a = text.find("$" + name)
        b = text.find("$End" + name)
        if a < 0 or b < 0:
            raise ValueError("%s: no $%s section" % (path, name))
        return text[a + len(name) + 1:b].split()

    fmt = section("MeshFormat")
    if not fmt[0].startswith("4") or int(fmt[1]) != 0:
        raise ValueError("%s: only ASCII MSH 4.x files are supported (got version %s, file-type %s)" % (path, fmt[0], fmt[1]))
    tok = section("Nodes")
    nblocks, nnodes, _, maxtag = (int(t) for t in tok[:4])
    coords = np.zeros((maxtag, 2))
    i = 4
    for _ in range(nblocks):
        _, _, parametric, nb = (int(t) for t in tok[i:i + 4])
        i += 4
        tags = np.array(tok[i:i + nb], dtype=np.int64)
        i += nb
        width = 3 + (int(tok[i - nb - 4]) if parametric else 0)     # x y z [+ u (v (w)) when parametric]
        xyz = np.array(tok[i:i + nb * width], dtype=np.float64).reshape(nb, width)
        i += nb * width
        coords[tags - 1] = xyz[:, :2]
    tok = section("Elements")
    nblocks = int(tok[0])
    i = 4
    conn, side_nodes = [], {int(s): set() for s in sides}
    for _ in range(nblocks):
        edim, etag, etype, nb = (int(t) for t in tok[i:i + 4])
        i += 4
        if etype not in _MSH_TYPES:
            raise ValueError("%s: unsupported MSH element type %d" % (path, etype))
        npe = _MSH_TYPES[etype][1]
        rows = np.array(tok[i:i + nb * (npe + 1)], dtype=np.uint64).reshape(nb, npe + 1)[:, 1:]
        i += nb * (npe + 1)
        if edim == 2 and (surface is None or etag == surface):
            conn.append(rows)
        elif edim == 1 and etag in side_nodes:
            side_nodes[etag].update(int(t) for t in rows.ravel())
    if not conn:
        raise ValueError("%s: no 2-D elements" % path)
    if len({c.shape[1] for c in conn}) != 1:
        raise ValueError("%s: mixed 2-D element types (the reference supports one type per mesh, elements.py:55)" % path)
    conectivity = [r for r in np.concatenate(conn)]
    return conectivity, coords, [side_nodes[int(s)] for s in sides]


def writeMsh(path, conectivity, coordinates, bcNodesList):
    """Writes the mesh as ASCII MSH 4.1 (nodes in one block on surface 1, boundary line elements on curves
    1..4 in the reference's side order, 2-D elements on surface 1) so that ``readMsh`` -- and gmsh -- read it."""
    conn = np.asarray(conectivity if isinstance(conectivity, np.ndarray) else np.stack([np.asarray(c) for c in conectivity]),
                      dtype=np.int64)
    xy = np.asarray(coordinates, dtype=np.float64)
    nnpe = conn.shape[1]
    etype = {3: 2, 4: 3, 9: 10}[nnpe]
    ncorner = 3 if nnpe == 3 else 4
    # boundary edges of the mesh: element sides whose corner pair occurs once
    sides = []
    for k in range(ncorner):
        a, b = conn[:, k], conn[:, (k + 1) % ncorner]
        mid = conn[:, ncorner + k] if nnpe == 9 else None
        sides.append((a, b, mid))
    a = np.concatenate([s[0] for s in sides])
    b = np.concatenate([s[1] for s in sides])
    mid = np.concatenate([s[2] for s in sides]) if nnpe == 9 else None
    key = np.minimum(a, b) * (xy.shape[0] + 1) + np.maximum(a, b)
    _, first, count = np.unique(key, return_index=True, return_counts=True)
    bnd = first[count == 1]
    lines = []
    for k, nodes in enumerate(bcNodesList, start=1):
        nodes = np.fromiter((int(t) for t in nodes), dtype=np.int64)
        on = np.isin(a[bnd], nodes) & np.isin(b[bnd], nodes)
        e = bnd[on]
        rows = np.stack([a[e], b[e]] + ([mid[e]] if nnpe == 9 else []), axis=1)
        lines.append((k, rows))
    nline = sum(len(r) for _, r in lines)
    with open(path, "w") as f:
        f.write("$MeshFormat\n4.1 0 8\n$EndMeshFormat\n")
        f.write("$Nodes\n1 %d 1 %d\n2 1 0 %d\n" % (xy.shape[0], xy.shape[0], xy.shape[0]))
        f.write("\n".join(str(t) for t in range(1, xy.shape[0] + 1)) + "\n")
        f.write("\n".join("%r %r 0" % (float(p[0]), float(p[1])) for p in xy) + "\n$EndNodes\n")
        f.write("$Elements\n%d %d 1 %d\n" % (len(lines) + 1, nline + len(conn), nline + len(conn)))
        tag = 1
        for k, rows in lines:
            f.write("1 %d %d %d\n" % (k, 8 if nnpe == 9 else 1, len(rows)))
            for r in rows:
                f.write("%d %s\n" % (tag, " ".join(str(int(t)) for t in r)))
                tag += 1
        f.write("2 1 %d %d\n" % (etype, len(conn)))
        for r in conn:
            f.write("%d %s\n" % (tag, " ".join(str(int(t)) for t in r)))
            tag += 1
        f.write("$EndElements\n")
