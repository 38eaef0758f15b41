#!/usr/bin/env python3
"""Times the UNMODIFIED reference (imported in place by oracle/ref_runner.py) on BASELINE configs[0]:
FemProblem(...) -> setMatrix -> setBoundaryConditions -> assemble -> solveProblem on the 64 x 64 Quad4 unit square,
one host core (the reference is single-threaded Python).  TEST / MEASUREMENT INFRASTRUCTURE: runs only where
/root/reference exists (the authoring container, not the GPU box); its output is committed under profiles/.

    python oracle/time_reference.py [n] > profiles/r02_reference_cfg1_cpu.json
"""
import contextlib
import io
import json
import os
import platform
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from feprogram_b200 import mesh  # noqa: E402
from oracle import ref_runner  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    conn, coords, bc = mesh.structured_quad4(n)
    reps = []
    for _ in range(3):
        fem = ref_runner.make_problem(conn, coords)
        with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            fem.setBoundaryConditions(bc)
            t0 = time.perf_counter()
            fem.assemble()
            t1 = time.perf_counter()
            U = fem.solveProblem()
            t2 = time.perf_counter()
        reps.append((t1 - t0, t2 - t1))
    asm = sorted(r[0] for r in reps)[1]
    sol = sorted(r[1] for r in reps)[1]
    import scipy
    print(json.dumps({
        "what": "unmodified reference (/root/reference/elements.py:165-200 assemble, :241-243 solveProblem), median of 3",
        "kind": "reference", "cores": 1, "where": "authoring container (no GPU); the reference cannot travel to the GPU box",
        "workload": "quad4_%dx%d_2dof_reference_operator (BASELINE configs[0])" % (n, n), "elements": int(conn.shape[0]),
        "assemble_s": asm, "solve_s": sol, "elements_assembled_per_s": conn.shape[0] / asm,
        "U_checksum": float(np.abs(U).sum()), "python": platform.python_version(), "numpy": np.__version__,
        "scipy": scipy.__version__, "host_cpu_count": os.cpu_count()}, indent=1))


if __name__ == "__main__":
    main()
