#!/usr/bin/env python3
"""Driver with the reference's call order and log lines (/root/reference/main.py:10-50).

    python -m feprogram_b200.main <outname> [--nx N --ny N [--quad9]]

Without options the mesh comes from ``gmsh_api.getMeshInfo()`` like the reference's (main.py:12; gmsh when
importable, else the equivalent structured 8 x 40 quad mesh of the 2 x 10 rectangle) and the result is
written with ``writeXML_nodeData`` (main.py:50).  With --nx/--ny a structured mesh of that size is used.
"""
import logging
import sys
from datetime import datetime

import numpy as np

from . import mesh
from .elements import FemProblem
from .gmsh_api import getMeshInfo
from .python_to_xml import writeXML, writeXML_nodeData  # noqa: F401  (the reference's import line, main.py:7)


class _Laps:
    """Wall-clock laps logged with the reference's labels (main.py:14-46)."""

    def __init__(self):
        self.start = self.last = datetime.now()

    def lap(self, label):
        now = datetime.now()
        logging.info("%s: %f sec", label, (now - self.last).total_seconds())
        self.last = now

    def total(self, label):
        logging.info("%s: %f sec", label, (datetime.now() - self.start).total_seconds())


def run(outname, nx=None, ny=None, quad9=False):
    logging.basicConfig(level="INFO")
    laps = _Laps()
    if nx is None:
        conectivity, coordinates, bcNodeTags = getMeshInfo()
    else:
        gen = mesh.structured_quad9 if quad9 else mesh.structured_quad4
        conectivity, coordinates, bcNodeTags = gen(nx, ny or nx)
        coordinates = coordinates * np.array([2.0, 10.0])   # the reference's 2 x 10 rectangle (gmsh_api.py:36-40)
    laps.lap("Crear malla")
    fem = FemProblem(conectivity, coordinates)
    laps.lap("Crear elemento fem")
    for label, call in (("Seteo inicial de matrices", fem.setMatrix),
                        ("Seteo de BC en nodos", lambda: fem.setBoundaryConditions(bcNodeTags)),   # prints both dof sets, as the reference
                        ("Ensamblaje de matrices", fem.assemble)):
        call()
        laps.lap(label)
    U = fem.solveProblem()
    laps.lap("Calculo de velocidades")
    U = U.reshape((coordinates.shape[0], 2))
    laps.total("Tiempo total")
    writeXML_nodeData(coordinates, conectivity, U, outname)
    return U


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("outname")
    ap.add_argument("--nx", type=int, default=None)
    ap.add_argument("--ny", type=int, default=None)
    ap.add_argument("--quad9", action="store_true")
    a = ap.parse_args()
    run(a.outname, a.nx, a.ny, a.quad9)
