#!/usr/bin/env python3
"""Generate tests/golden/*.npz by executing the unmodified reference in this container.

    python oracle/make_golden.py

The reference ships no golden vectors, so these fixtures are outputs of the
reference itself run here (numpy / scipy versions recorded inside each file); they are what
pins the oracle and the CUDA path on the GPU box, where /root/reference does not exist.
"""
import os
import sys

import numpy as np
import scipy

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from feprogram_b200 import mesh  # noqa: E402  (host-side generators only; no CUDA involved)
from oracle import ref_runner  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def case_mesh(kind, n, perturbed, shuffled):
    if kind == "u4":            # genuinely unstructured all-quad mesh (node degrees 4..15, random numbering)
        return mesh.unstructured_quad4(n, seed=5)
    if kind == 4:
        conn, coords, bc = mesh.structured_quad4(n)
        mx = n
    else:
        conn, coords, bc = mesh.structured_quad9(n)
        mx = 2 * n
    if perturbed:
        coords = mesh.perturb_interior(coords, mx, mx, 0.2, seed=0)
    if shuffled:
        conn, coords, bc, _ = mesh.shuffle_numbering(conn, coords, bc, seed=1)
    return conn, coords, bc


def bc_arrays(bc):
    return {"bc%d" % i: np.array(sorted(s), dtype=np.int64) for i, s in enumerate(bc)}


def full_case(name, kind, n, perturbed, shuffled, nke=6):
    conn, coords, bc = case_mesh(kind, n, perturbed, shuffled)
    K0 = ref_runner.assemble_prebc(conn, coords)
    run = ref_runner.full_run(conn, coords, bc)
    fem = ref_runner.make_problem(conn, coords)
    ke = np.array([ref_runner.element_matrix(fem, conn[e] - 1) for e in range(min(nke, len(conn)))])
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        conn=conn, coords=coords, **bc_arrays(bc),
        k0_indptr=K0.indptr, k0_indices=K0.indices, k0_data=K0.data,
        k_indptr=run["K"].indptr, k_indices=run["K"].indices, k_data=run["K"].data,
        brhs=run["brhs"], U=run["U"], dofDir=run["dofDir"], dofNeu=run["dofNeu"], ke=ke,
        versions=np.array([np.__version__, scipy.__version__]),
    )
    print(name, "nnz0", K0.nnz, "nnz", run["K"].nnz, "|U|max", np.abs(run["U"]).max())


def scalar_case(kind, n):
    conn, coords, bc = case_mesh(kind, n, False, False)
    K0 = ref_runner.assemble_prebc(conn, coords)
    run = ref_runner.full_run(conn, coords, bc)
    U = run["U"]
    return np.array([K0.nnz, np.sqrt((K0.data ** 2).sum()), K0.diagonal().sum(), run["K"].nnz,
                     len(run["dofDir"]), len(run["dofNeu"]), U.min(), U.max(), np.linalg.norm(U)])


def mass_golden():
    """N^T N element matrices built from the reference's own H / gpWei / Hrs objects (ref_runner.mass_matrix)."""
    X4 = np.array([[0, 0], [2, 0.1], [2.3, 1.2], [-0.1, 0.9]], dtype=np.float64)
    fem4 = ref_runner.make_problem(np.array([[1, 2, 3, 4]]), X4)
    rng = np.random.default_rng(7)
    c9, x9, _ = mesh.structured_quad9(1)
    x9 = 3.0 * x9 + rng.uniform(-0.12, 0.12, size=x9.shape)
    fem9 = ref_runner.make_problem(c9, x9)
    conn, coords, _ = case_mesh(4, 6, True, True)
    fem = ref_runner.make_problem(conn, coords)
    me = np.array([ref_runner.mass_matrix(fem, conn[e] - 1) for e in range(len(conn))])
    np.savez_compressed(os.path.join(OUT, "mass_single.npz"), X4=X4, me4=ref_runner.mass_matrix(fem4, np.arange(4)),
                        conn9=c9, X9=x9, me9=ref_runner.mass_matrix(fem9, c9[0] - 1), conn=conn, coords=coords, me=me,
                        versions=np.array([np.__version__, scipy.__version__]))
    print("mass_single.npz written")


def post_golden():
    """Velocity gradients through the reference's own Elem.getHgrad on small perturbed, shuffled meshes."""
    out = {}
    rng = np.random.default_rng(3)
    for kind, n in ((4, 5), (9, 3)):
        conn, coords, _ = case_mesh(kind, n, True, True)
        U = rng.standard_normal(2 * coords.shape[0])
        fem = ref_runner.make_problem(conn, coords)
        out.update({"conn%d" % kind: conn, "coords%d" % kind: coords, "U%d" % kind: U,
                    "grad%d" % kind: ref_runner.gauss_gradients(fem, conn, U)})
    np.savez_compressed(os.path.join(OUT, "post_gradients.npz"), versions=np.array([np.__version__, scipy.__version__]),
                        **out)
    print("post_gradients.npz written")


def main():
    os.makedirs(OUT, exist_ok=True)
    if sys.argv[1:] == ["unstructured"]:
        return full_case("q4_unstructured", "u4", 6, False, False)
    if sys.argv[1:] == ["mass"]:
        return mass_golden()
    if sys.argv[1:] == ["post"]:
        return post_golden()
    # tables (T1-T4)
    d4, h4, w4 = ref_runner.tables(4)
    d9, h9, w9 = ref_runner.tables(9)
    # single distorted elements (E1-E3)
    X4 = np.array([[0, 0], [2, 0.1], [2.3, 1.2], [-0.1, 0.9]], dtype=np.float64)
    fem4 = ref_runner.make_problem(np.array([[1, 2, 3, 4]]), X4)
    ke4 = ref_runner.element_matrix(fem4, np.arange(4))
    rng = np.random.default_rng(7)
    c9, x9, _ = mesh.structured_quad9(1)
    x9 = 3.0 * x9 + rng.uniform(-0.12, 0.12, size=x9.shape)
    fem9 = ref_runner.make_problem(c9, x9)
    ke9 = ref_runner.element_matrix(fem9, c9[0] - 1)
    np.savez_compressed(os.path.join(OUT, "tables_and_single.npz"), dH4=d4, H4=h4, w4=w4, dH9=d9, H9=h9,
                        w9=w9, X4=X4, ke4=ke4, conn9=c9, X9=x9, ke9=ke9,
                        versions=np.array([np.__version__, scipy.__version__]))
    # full small problems: structure, values, BCs, solution
    full_case("q4_4x4", 4, 4, False, False)
    full_case("q4_12x12_pert_shuf", 4, 12, True, True)
    full_case("q4_16x16_pert", 4, 16, True, False)
    full_case("q9_4x4", 9, 4, False, False)
    full_case("q9_6x6_pert_shuf", 9, 6, True, True)
    full_case("q4_unstructured", "u4", 6, False, False)
    # scalar known answers on BASELINE cfg1 and a Quad9 case
    np.savez_compressed(os.path.join(OUT, "scalars.npz"),
                        fields=np.array(["nnz0", "normF0", "trace0", "nnz_bc", "nDir", "nNeu", "Umin", "Umax",
                                         "Unorm"]),
                        q4_64=scalar_case(4, 64), q9_32=scalar_case(9, 32), q4_4=scalar_case(4, 4),
                        q9_4=scalar_case(9, 4),
                        versions=np.array([np.__version__, scipy.__version__]))
    # .vtu writer (python_to_xml.py:74-141): the reference's own output on two tiny meshes
    import importlib.util
    spec = importlib.util.spec_from_file_location("_ref_writer", os.path.join(ref_runner.REF_ROOT, "python_to_xml.py"))
    writer = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(writer)
    rng = np.random.default_rng(0)
    for k, gen in enumerate((mesh.structured_quad4, mesh.structured_quad9)):
        conn, coords, _ = gen(3, 2)
        U = rng.standard_normal((coords.shape[0], 2)) * 1e3
        writer.writeXML_nodeData(coords * np.array([2.0, 10.0]), conn, U, os.path.join(OUT, "writer_case%d" % k))
    mass_golden()
    post_golden()
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
