"""``getMeshInfo()`` with the reference's contract (/root/reference/gmsh_api.py:13-113; next-row f-2 of
SURVEY.md §8, not on the hot path): returns ``(conectivity, coordinates, bcNodesList)`` of the 2 x 10
rectangle meshed with quads of size lc = 0.25 -- 1-based tags, coordinates (nnode, 3), four boundary tag
sets ordered [bottom, right, top, left].

The reference drives the gmsh SDK (an external C++ mesher).  When ``gmsh`` is importable the same model is
built with it; otherwise (this image has no gmsh) the transfinite-equivalent structured 8 x 40 quad mesh
of the same rectangle is generated, which is what gmsh's recombined mesh of this geometry amounts to.
"""
import numpy as np

from . import mesh

LX, LY, LC = 2.0, 10.0, 0.25


def _structured():
    nx, ny = int(round(LX / LC)), int(round(LY / LC))
    conn, xy, bc = mesh.structured_quad4(nx, ny, dtype=np.uint64)
    coords = np.zeros((xy.shape[0], 3))
    coords[:, 0] = xy[:, 0] * LX
    coords[:, 1] = xy[:, 1] * LY
    return list(conn), coords, bc


def getMeshInfo():
    try:
        import gmsh  # noqa: F401
    except Exception:
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
        gmsh.model.geo.synchronize()
        gmsh.option.setNumber("Mesh.RecombineAll", 1)
        gmsh.model.mesh.generate(2)
        tags, xyz, _ = gmsh.model.mesh.getNodes()
        coords = np.zeros((int(tags.max()), 3))
        coords[np.asarray(tags, dtype=np.int64) - 1] = np.asarray(xyz).reshape(-1, 3)
        _, _, enodes = gmsh.model.mesh.getElements(2)
        conn = [np.asarray(r) for r in np.asarray(enodes[0], dtype=np.uint64).reshape(-1, 4)]
        sides = []
        for line in (1, 2, 3, 4):          # bottom, right, top, left
            ntags, _, _ = gmsh.model.mesh.getNodes(1, line, includeBoundary=True)
            sides.append(set(int(t) for t in ntags))
        return conn, coords, sides
    finally:
        gmsh.finalize()
