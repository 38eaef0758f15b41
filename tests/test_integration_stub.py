"""The ctypes binding INTEGRATION.md shows a reference maintainer (section 2) is executed as written: the code block is
extracted from the document, pointed at the in-tree library, and its assemble() / solveProblem() are run on a frozen
reference case.  Keeps the documented boundary from rotting."""
import os
import re
import types

import numpy as np
import pytest

from conftest import ROOT, bc_sets_of

pytestmark = pytest.mark.gpu


def _stub_source():
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    src = next(b for b in blocks if "class FemProblem" in b and "fe_pcg" in b)
    lib = os.path.join(ROOT, "feprogram_b200", "libfeprogram_b200.so")
    assert 'C.CDLL("libfeprogram_b200.so")' in src
    return src.replace('C.CDLL("libfeprogram_b200.so")', "C.CDLL(%r)" % lib)


@pytest.mark.parametrize("name", ["q4_12x12_pert_shuf", "q4_unstructured"])
def test_documented_ctypes_binding_reproduces_the_reference(golden, name):
    import torch
    from feprogram_b200 import build
    build.build()
    import feprogram_b200.ops  # noqa: F401  (loads torch's libnccl before the bare CDLL of the stub)
    assert torch.cuda.is_available()
    z = golden(name)
    ns = {}
    exec(compile(_stub_source(), "INTEGRATION.md", "exec"), ns)
    fem = object.__new__(ns["FemProblem"])
    fem.conectivity = [np.asarray(r, dtype=np.uint64) for r in z["conn"]]      # what gmsh hands out
    fem.coordinates = np.concatenate([z["coords"], np.zeros((z["coords"].shape[0], 1))], axis=1)   # (nnode, 3)
    fem.nelem, fem.nnode = len(z["conn"]), z["coords"].shape[0]
    fem.elem = types.SimpleNamespace(nnodesloc=z["conn"].shape[1])
    fem.dofDir, fem.dofNeu = z["dofDir"].tolist(), z["dofNeu"].tolist()
    fem.assemble()
    rp, ci, vals, rhs, mask = fem._gpu
    assert np.array_equal(rp.cpu().numpy(), z["k0_indptr"]) and np.array_equal(ci.cpu().numpy(), z["k0_indices"])
    assert np.abs(vals.cpu().numpy() - z["k0_data"]).max() <= 1e-12 * np.abs(z["k0_data"]).max()
    assert np.array_equal(rhs.cpu().numpy(), z["brhs"])
    U = fem.solveProblem()
    assert np.linalg.norm(U - z["U"]) / np.linalg.norm(z["U"]) < 1e-8
    _ = bc_sets_of
