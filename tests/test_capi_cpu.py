"""CPU-side checks of the C-ABI library: it loads, exports every symbol the header declares, and its
host-only entry points (tables, limits, error reporting) behave.  No compute calls (no GPU here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def lib():
    from feprogram_b200 import build, _lib
    build.build()
    return _lib.lib()


def header_functions():
    text = open(os.path.join(ROOT, "include", "feprogram_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fe_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    from feprogram_b200 import _lib
    names = header_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), "library does not export %s" % n
    assert sorted(_lib.PROTOTYPES) == names, "ctypes prototypes and header disagree"


def test_version_and_limits(lib):
    from feprogram_b200 import _lib
    assert "sm_100a" in _lib.version()
    lim = _lib.limits()
    assert lim["max_adj"] >= 25 and lim["max_nnpe"] == 9


def test_tables_bit_exact_against_reference_golden(lib, golden):
    from feprogram_b200 import ops
    z = golden("tables_and_single")
    for et, tag in ((4, "4"), (9, "9")):
        dH, H, w = ops.tables(et)
        assert np.array_equal(dH, z["dH" + tag])
        assert np.array_equal(H, z["H" + tag])
        assert np.array_equal(w, z["w" + tag])


def test_unknown_element_type_reports_error(lib):
    from feprogram_b200 import _lib
    rc = lib.fe_tables(5, None, None, None, None, None)
    assert rc == -2
    assert "Invalid elemType" in _lib.last_error()


def test_null_arguments_are_rejected_without_touching_the_gpu(lib):
    from feprogram_b200 import _lib
    rc = lib.fe_spmv(10, 180, 0, None, None, None, None, None, None, None)
    assert rc == -1 and "null" in _lib.last_error()
    rc = lib.fe_assemble(4, 7, 1, 1, None, None, None, None, None, None, 9, None, None)
    assert rc == -1


def test_facade_host_side_matches_reference_api(golden):
    """Host-only parts of the drop-in classes: tables, BC dof sets, getRowData, Elem helpers."""
    import contextlib
    import io
    from feprogram_b200.elements import Elem, FemProblem
    z = golden("q4_4x4")
    fem = FemProblem(z["conn"], z["coords"])
    assert (fem.nnode, fem.nelem, fem.elem.elemType, fem.elem.nnodesloc) == (25, 16, "Quad4", 4)
    fem.setMatrix()
    t = golden("tables_and_single")
    assert np.array_equal(np.array([np.asarray(h) for h in fem.Hrs]), t["dH4"])
    assert np.array_equal(np.array([np.asarray(h).ravel() for h in fem.H]), t["H4"])
    assert np.array_equal(np.array(fem.gpWei, dtype=float), t["w4"])
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        fem.setBoundaryConditions([set(z["bc%d" % i].tolist()) for i in range(4)])
    assert buf.getvalue().count("{") == 2          # the reference prints both dof sets
    assert sorted(fem.dofDir) == z["dofDir"].tolist() and sorted(fem.dofNeu) == z["dofNeu"].tolist()
    row, col = fem.getRowData(z["conn"][0])
    d = [(int(t_) - 1) * 2 + c for t_ in z["conn"][0] for c in (0, 1)]
    assert col[:8] == d and row[:8] == [d[0]] * 8 and len(row) == 64
    # Elem.funB against the closed form B of SURVEY.md §8(a) E2
    e = Elem(4, "Quad4")
    X = e.getpos([0, 1, 6, 5], z["coords"])
    J = fem.Hrs[0] * X
    B = np.asarray(e.funB(fem.Hrs[0], J))
    Dh = np.asarray(np.linalg.inv(J) * fem.Hrs[0])
    assert np.allclose(B[0, 0::2], Dh[0]) and np.allclose(B[1, 1::2], Dh[1])
    assert np.allclose(B[2, 0::2], Dh[1]) and np.allclose(B[2, 1::2], Dh[0]) and not B[3].any()
    with pytest.raises(Exception, match="Invalid elemType"):
        fem.elem.elemType = "Hex8"
        fem.setMatrix()


def test_hot_path_fails_loudly_without_cuda(golden):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from feprogram_b200.elements import FemProblem
    z = golden("q4_4x4")
    fem = FemProblem(z["conn"], z["coords"])
    fem.setMatrix()
    fem.setBoundaryConditions([set(z["bc%d" % i].tolist()) for i in range(4)], verbose=False)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        fem.assemble()


def test_ctypes_prototypes_match_header_arity_and_kinds(lib):
    """Each ctypes prototype has the parameter count and pointer/scalar kinds the header declares."""
    import ctypes as C
    from feprogram_b200 import _lib
    text = open(os.path.join(ROOT, "include", "feprogram_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    decls = re.findall(r"\b(fe_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S)
    assert len(decls) == len(_lib.PROTOTYPES)
    for name, params in decls:
        params = " ".join(params.split())
        plist = [] if params in ("", "void") else [p.strip() for p in params.split(",")]
        _, argtypes = _lib.PROTOTYPES[name]
        assert len(plist) == len(argtypes), "%s: header has %d parameters, ctypes %d" % (name, len(plist), len(argtypes))
        for p, a in zip(plist, argtypes):
            is_ptr = "*" in p or "[" in p
            a_ptr = a in (C.c_void_p, C.c_char_p) or hasattr(a, "contents")
            assert is_ptr == a_ptr, "%s: parameter '%s' vs ctypes %s" % (name, p, a)
            if not is_ptr:
                want = {"int": C.c_int, "int32_t": C.c_int32, "int64_t": C.c_int64, "double": C.c_double,
                        "size_t": C.c_size_t}[p.split()[-2] if len(p.split()) > 1 else p]
                assert a is want or (want is C.c_int and a is C.c_int32), "%s: '%s' vs %s" % (name, p, a)
