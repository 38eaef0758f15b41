"""Runs the UNMODIFIED reference (/root/reference/elements.py) on synthetic meshes.

TEST INFRASTRUCTURE.  /root/reference exists only in the authoring container; on the GPU box
``available()`` is False and the tests fall back to the frozen outputs in ``tests/golden/``
(written by ``oracle/make_golden.py`` from this module).  Nothing is copied from the
reference: it is imported in place, read-only, under a private module name so that it never
collides with the drop-in ``feprogram_b200.elements``.
"""
from __future__ import annotations

import contextlib
import importlib.util
import io
import os
import sys
import warnings

import numpy as np

REF_ROOT = os.environ.get("FEPROGRAM_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "elements.py"))


_mod = None


def ref_elements():
    """The reference's ``elements`` module (with its ``tools.interpFunctions``), imported in place."""
    global _mod
    if _mod is not None:
        return _mod
    if not available():
        raise RuntimeError("reference not present at %s" % REF_ROOT)
    saved_path = list(sys.path)
    saved_tools = {k: v for k, v in sys.modules.items() if k == "tools" or k.startswith("tools.")}
    for k in saved_tools:
        del sys.modules[k]
    sys.path.insert(0, REF_ROOT)
    try:
        spec = importlib.util.spec_from_file_location("_feprogram_reference_elements",
                                                      os.path.join(REF_ROOT, "elements.py"))
        mod = importlib.util.module_from_spec(spec)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            spec.loader.exec_module(mod)
    finally:
        sys.path[:] = saved_path
        for k in [k for k in sys.modules if k == "tools" or k.startswith("tools.")]:
            sys.modules["_feprogram_reference_" + k] = sys.modules.pop(k)
        sys.modules.update(saved_tools)
    _mod = mod
    return mod


def make_problem(conn1, coords):
    """FemProblem(conectivity, coordinates) + setMatrix()  (main.py:17-22)."""
    m = ref_elements()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        fem = m.FemProblem(np.asarray(conn1), np.asarray(coords, dtype=np.float64))
        fem.setMatrix()
    return fem


def tables(nnpe: int):
    conn = np.arange(1, nnpe + 1)[None, :]
    fem = make_problem(conn, np.zeros((nnpe, 2)))
    return (np.array([np.asarray(h) for h in fem.Hrs]), np.array([np.asarray(h).ravel() for h in fem.H]),
            np.array(fem.gpWei, dtype=np.float64))


def element_matrix(fem, tags0):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return np.asarray(fem.getKe(np.asarray(tags0)))


def mass_matrix(fem, tags0):
    """The N^T N product of the north_star built from the reference's OWN objects: fem.H / fem.gpWei / fem.Hrs
    (elements.py:94-140; the reference fills them and never uses H or gpWei), Elem.getpos and np.linalg.det as
    getKe does (elements.py:159-160): Me = sum_g gpWei_g H_g^T H_g detJ_g on each of the two components."""
    tags0 = np.asarray(tags0)
    X = fem.elem.getpos(tags0, fem.coordinates)
    n = len(tags0)
    me = np.zeros((2 * n, 2 * n))
    for g in range(len(fem.Hrs)):
        J = fem.Hrs[g] * X
        h = np.asarray(fem.H[g]).ravel()
        m = fem.gpWei[g] * np.outer(h, h) * np.linalg.det(J)
        me[0::2, 0::2] += m
        me[1::2, 1::2] += m
    return me


def gauss_gradients(fem, conn1, U):
    """H_grad u_e at every Gauss point through the reference's own Elem.getHgrad (elements.py:25-32)."""
    U = np.asarray(U, dtype=np.float64).reshape(-1)
    out = []
    for tags in np.asarray(conn1):
        t0 = np.asarray(tags) - 1
        X = fem.elem.getpos(t0, fem.coordinates)
        ue = np.matrix(np.stack([U[2 * t0], U[2 * t0 + 1]], axis=1).ravel()).T
        out.append([np.asarray(fem.elem.getHgrad(fem.Hrs[g], fem.Hrs[g] * X) * ue).ravel() for g in range(len(fem.Hrs))])
    return np.array(out)


def assemble_prebc(conn1, coords):
    """Pre-BC K straight from the reference: dofDir = dofNeu = [] then assemble()."""
    fem = make_problem(conn1, coords)
    fem.dofDir, fem.dofNeu = [], []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        fem.assemble()
    K = fem.K.tocsr()
    K.sort_indices()
    return K


def full_run(conn1, coords, bc_sets):
    """ctor -> setMatrix -> setBoundaryConditions -> assemble -> solveProblem (main.py:17-40)."""
    fem = make_problem(conn1, coords)
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        fem.setBoundaryConditions(bc_sets)        # prints both sets, elements.py:226-227
        fem.assemble()
        U = fem.solveProblem()
    K = fem.K.tocsr()
    K.sort_indices()
    b = np.asarray(fem.brhs.todense()).ravel()
    return dict(K=K, brhs=b, U=np.asarray(U), dofDir=np.sort(np.asarray(fem.dofDir, dtype=np.int64)),
                dofNeu=np.sort(np.asarray(fem.dofNeu, dtype=np.int64)))
