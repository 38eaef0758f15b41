#!/usr/bin/env python3
"""Stage-by-stage wall-clock of the public-API assembly (host buffers -> K, rhs) -- developer tool.
Synchronises between stages, so the sum is an upper bound of the pipelined e2e time bench.py reports."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    import torch
    from feprogram_b200 import build
    build.build()
    from feprogram_b200 import mesh, ops
    from feprogram_b200.elements import FemProblem
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    dev = torch.device("cuda", 0)
    conn_h, coords_h, bc = mesh.structured_quad4(n, dtype=np.int32)
    d_dofs, n_dofs = [np.unique(np.concatenate([2 * (np.array(sorted(s), dtype=np.int64) - 1),
                                                 2 * (np.array(sorted(s), dtype=np.int64) - 1) + 1])) for s in (bc[0], bc[3])]
    conn1 = torch.from_numpy(conn_h).pin_memory()
    xy_p = torch.from_numpy(coords_h).pin_memory()
    out_rhs = torch.empty(2 * coords_h.shape[0], dtype=torch.float64).pin_memory()
    sync = torch.cuda.synchronize

    def lap(t, name, rec):
        sync()
        now = time.perf_counter()
        rec[name] = round(1e3 * (now - t), 3)
        return now

    rows = []
    for rep in range(4):
        rec = {}
        sync()
        t = t00 = time.perf_counter()
        fem = FemProblem(conn1.numpy(), xy_p.numpy(), device=dev)
        fem.setMatrix()
        fem.dofDir, fem.dofNeu = d_dofs, n_dofs
        st = fem._state()
        t = lap(t, "upload+tags", rec)
        pat = ops.symbolic(st["conn_s"], fem.nnode, 4, 2)
        t = lap(t, "symbolic", rec)
        plan = ops.tile_plan(st["conn_s"], st["coords_s"], pat)
        t = lap(t, "tile_plan", rec)
        st["pattern"], st["plan"] = pat, plan
        fem.assemble()
        t = lap(t, "assemble(numeric+bc)", rec)
        chk = st["vals"].sum().item()
        b = fem.rhs_host()
        t = lap(t, "d2h+checksum", rec)
        rec["total"] = round(1e3 * (t - t00), 3)
        del fem, st, pat, plan
        t = lap(t, "free", rec)
        rows.append(rec)
    # does a concurrent host -> device copy slow the symbolic kernels down?  (same stage, 268 MB upload in flight)
    side = torch.cuda.Stream(dev)
    for rep in range(2):
        fem = FemProblem(conn1.numpy(), xy_p.numpy(), device=dev)
        fem.setMatrix()
        st = fem._state()
        sync()
        t = time.perf_counter()
        with torch.cuda.stream(side):
            dummy = xy_p.to(dev, non_blocking=True)
        pat = ops.symbolic(st["conn_s"], fem.nnode, 4, 2, csr="lazy")
        torch.cuda.current_stream(dev).synchronize()
        t1 = time.perf_counter()
        sync()
        t2 = time.perf_counter()
        pat2 = ops.symbolic(st["conn_s"], fem.nnode, 4, 2, csr="lazy")
        sync()
        t3 = time.perf_counter()
        rows.append({"symbolic_lazy_with_h2d": round(1e3 * (t1 - t), 3), "h2d_total": round(1e3 * (t2 - t), 3),
                     "symbolic_lazy_alone": round(1e3 * (t3 - t2), 3)})
        del fem, st, pat, pat2, dummy
    # the pipelined call, as bench.py times it
    for rep in range(3):
        sync()
        t0 = time.perf_counter()
        fem = FemProblem(conn1.numpy(), xy_p.numpy(), device=dev, trace=(rep == 2))
        fem.setMatrix()
        fem.dofDir, fem.dofNeu = d_dofs, n_dofs
        fem.assemble()
        chk = fem._dev["vals"].sum().item()
        b = fem.rhs_host()
        sync()
        rows.append({"pipelined_ms": round(1e3 * (time.perf_counter() - t0), 3), "alloc": torch.cuda.memory_stats()["num_device_alloc"],
                     "stages": fem.stage_times()})
        del fem
    print(json.dumps(rows, indent=1))


if __name__ == "__main__":
    main()
