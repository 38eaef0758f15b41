"""ASCII .vtu writer with the reference's file format (next-row f-3 of SURVEY.md §8; not on the hot path).

``writeXML_nodeData(coords, conect, nodedata, nameFile)`` produces byte for byte what
/root/reference/python_to_xml.py:74-141 writes -- PointData vector "Velocidad" (2 components), Points with
a 0.0 z, 0-based connectivity, offsets, VTK cell types 9 (quad) / 28 (biquadratic quad) -- but builds each
section with vectorised string operations and one write per section instead of one ``write`` per number.
"""
import numpy as np

__all__ = ["writeXML", "writeXML_nodeData", "writeVTK"]


def _s(a):
    """Shortest round-trip decimal strings, identical to ``str(np.float64)`` / ``'{}'.format``."""
    return np.asarray(a).astype(str)


def writeXML_nodeData(coords, conect, nodedata, nameFile, nodeStress=None, _name="Velocidad"):
    """``nodeStress`` (optional, (nnode, 3)): also writes the "Tensiones" PointData array in the layout of the
    block the reference keeps commented out (python_to_xml.py:19-22); None reproduces the reference's file."""
    coords = np.asarray(coords)
    nnode = coords.shape[0]
    nelem = len(conect)
    conn = np.asarray(conect)
    if conn.dtype == object:
        conn = np.stack([np.asarray(c) for c in conect])
    nnpe = conn.shape[1]
    cellType = 9 if nnpe == 4 else 28
    data = np.asarray(nodedata)
    with open(nameFile.split(".")[0] + ".vtu", "w") as f:
        f.write('<?xml version="1.0"?>\n<VTKFile type="UnstructuredGrid" version="0.1" byte_order="LittleEndian">\n')
        f.write(" <UnstructuredGrid>\n")
        f.write('  <Piece NumberOfPoints="{}" NumberOfCells="{}">\n'.format(nnode, nelem))
        f.write('   <PointData Vectors="Tensiones">\n')
        if nodeStress is not None:
            t = _s(np.asarray(nodeStress)[:, :3])
            f.write('    <DataArray Name="Tensiones" NumberOfComponents="3" type="Float32"  format="ascii">\n    ')
            f.write("".join(np.char.add(np.char.add(np.char.add(np.char.add(" ", t[:, 0]), np.char.add(" ", t[:, 1])),
                                                    np.char.add(" ", t[:, 2])), "\n").tolist()))
            f.write("\n    </DataArray>")
        f.write('    <DataArray Name="%s" NumberOfComponents="2" type="Float32" format="ascii">\n    ' % _name)
        d = _s(data[:, :2])
        f.write("".join(np.char.add(np.char.add(np.char.add(" ", d[:, 0]), np.char.add(" ", d[:, 1])), "\n").tolist()))
        f.write("\n    </DataArray>")
        f.write("\n   </PointData>\n")
        f.write("   <Points>\n")
        f.write('    <DataArray NumberOfComponents="3" type="Float32"  format="ascii">\n    ')
        c = _s(coords[:, :2])
        f.write("".join(np.char.add(np.char.add(np.char.add(" ", c[:, 0]), np.char.add(" ", c[:, 1])), " 0.0").tolist()))
        f.write("\n    </DataArray>")
        f.write("\n   </Points>\n")
        f.write("   <Cells>\n")
        f.write('    <DataArray type="Int32" Name="connectivity" format="ascii" RangeMin="" RangeMax="">\n')
        f.write("".join(np.char.add(" ", (conn.astype(np.int64) - 1).ravel().astype(str)).tolist()))
        f.write("\n    </DataArray>\n")
        f.write('    <DataArray type="Int32" Name="offsets" format="ascii" RangeMin="" RangeMax="">\n')
        f.write("".join(np.char.add(" ", (nnpe * np.arange(1, nelem + 1, dtype=np.int64)).astype(str)).tolist()))
        f.write("\n    </DataArray>\n")
        f.write('    <DataArray type="Int32" Name="types" format="ascii" RangeMin="" RangeMax="">\n')
        f.write((" " + str(cellType)) * nelem)
        f.write("\n    </DataArray>")
        f.write("\n   </Cells>\n")
        f.write("  </Piece>\n")
        f.write(" </UnstructuredGrid>\n")
        f.write("</VTKFile>\n")


def writeXML(coords, conect, nodedata, nameFile, nodeStress):
    """The reference's other .vtu writer (python_to_xml.py:1-69; main.py:7 imports it): same file as
    ``writeXML_nodeData`` with the point vector named "Displacement".  ``nodeStress`` is accepted and ignored,
    as in the reference (its "Tensiones" block is commented out there)."""
    writeXML_nodeData(coords, conect, nodedata, nameFile, None, _name="Displacement")


def writeVTK(coords, conect, nodedata, nameFile, nodeStress):
    """Legacy ASCII .vtk writer with the reference's layout (python_to_xml.py:145-180): 3-column coordinates,
    (nelem, 4) connectivity written as given, cell type 4 for every cell and the first component of
    ``nodedata`` as the scalar "Displacement" -- quirks kept, so the files compare equal."""
    coords, conn, data = np.asarray(coords), np.asarray(conect), np.asarray(nodedata)
    nnode, nelem = coords.shape[0], conn.shape[0]
    c = _s(coords[:, :3])
    rows = np.char.add(np.char.add(np.char.add(np.char.add(c[:, 0], " "), np.char.add(c[:, 1], " ")), c[:, 2]), "  \n")
    k = conn[:, :4].astype(np.int64).astype(str)
    cells = np.char.add("4 ", np.char.add(np.char.add(np.char.add(k[:, 0], " "), np.char.add(k[:, 1], " ")),
                                         np.char.add(np.char.add(k[:, 2], " "), np.char.add(k[:, 3], " \n"))))
    with open(nameFile.split(".")[0] + ".vtk", "w") as f:
        f.write("# vtk DataFile Version 2.0\nASCII\nDATASET UNSTRUCTURED_GRID\n \n")
        f.write("POINT {} float\n".format(nnode) + "".join(rows.tolist()) + "\n")
        f.write("CELLS {} {} \n".format(nelem, nelem * 5) + "".join(cells.tolist()) + "\n")
        f.write("CELL_TYPES {}\n".format(nelem) + "4\n" * nelem + "\n")
        f.write("POINT_DATA {}\nSCALARS Displacement float\nLOOKUP_TABLE default\n".format(nnode))
        f.write("".join(np.char.add(_s(data[:, 0]), "\n").tolist()) + "\n")
