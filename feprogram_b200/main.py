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
from .python_to_xml import writeXML_nodeData


def run(outname, nx=None, ny=None, quad9=False):
    t1 = datetime.now()
    logging.basicConfig(level="INFO")
    if nx is None:
        conectivity, coordinates, bcNodeTags = getMeshInfo()
    else:
        gen = mesh.structured_quad9 if quad9 else mesh.structured_quad4
        conectivity, coordinates, bcNodeTags = gen(nx, ny or nx)
        coordinates = coordinates * np.array([2.0, 10.0])   # the reference's 2 x 10 rectangle (gmsh_api.py:36-40)
    t2 = datetime.now()
    logging.info("Crear malla: %f sec", (t2 - t1).total_seconds())

    fem = FemProblem(conectivity, coordinates)
    t3 = datetime.now()
    logging.info("Crear elemento fem: %f sec", (t3 - t2).total_seconds())

    fem.setMatrix()
    t4 = datetime.now()
    logging.info("Seteo inicial de matrices: %f sec", (t4 - t3).total_seconds())

    fem.setBoundaryConditions(bcNodeTags, verbose=False)
    t5 = datetime.now()
    logging.info("Seteo de BC en nodos: %f sec", (t5 - t4).total_seconds())

    fem.assemble()
    t6 = datetime.now()
    logging.info("Ensamblaje de matrices: %f sec", (t6 - t5).total_seconds())

    U = fem.solveProblem()
    t7 = datetime.now()
    logging.info("Calculo de velocidades: %f sec", (t7 - t6).total_seconds())

    U = U.reshape((coordinates.shape[0], 2))
    t8 = datetime.now()
    logging.info("Tiempo total: %f sec", (t8 - t1).total_seconds())
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
