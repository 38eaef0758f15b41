#!/usr/bin/env python3
"""Times the numeric assembly kernels against each other on one mesh and checks them bit for bit.

    python tools/experiments/exp_ws.py KIND N [perturb] [reps]

Prints one JSON line.  Developer tool (not part of the product or the tests)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    import torch
    from feprogram_b200 import _lib
    if os.environ.get("FE_LIB"):                 # developer builds with extra -D flags (tools/experiments/variants)
        _lib.LIB_PATH = os.path.join(ROOT, os.environ["FE_LIB"])
    else:
        from feprogram_b200 import build
        build.build()
    from feprogram_b200 import mesh, ops
    kind, n = int(sys.argv[1]), int(sys.argv[2])
    perturb = len(sys.argv) > 3 and sys.argv[3] == "1"
    reps = int(sys.argv[4]) if len(sys.argv) > 4 else 10
    gen = {4: mesh.structured_quad4, 9: mesh.structured_quad9, 3: mesh.structured_tri3}[kind]
    conn, coords, _ = gen(n, dtype=np.int32)
    mx = 2 * n if kind == 9 else n
    if perturb:
        coords = mesh.perturb_interior(coords, mx, mx, 0.2, seed=7)
    conn0 = torch.from_numpy(conn - 1).cuda()
    xy = torch.from_numpy(coords).cuda()
    nnode = coords.shape[0]
    pat = ops.symbolic(conn0, nnode, kind, 2)
    out = {"kind": kind, "n": n, "perturb": perturb, "nelem": int(conn0.shape[0]), "nnz": int(pat.nnz)}
    bytes_alg = conn0.shape[0] * kind * 4 + nnode * 16 + pat.nnz * 8
    v_rows = torch.empty(pat.nnz, dtype=torch.float64, device="cuda")
    ops.assemble_values(conn0, xy, pat, None, v_rows)
    torch.cuda.synchronize()
    kernels = os.environ.get("FE_KERNELS", "ws,phased").split(",")
    out["lib"] = os.environ.get("FE_LIB", "release")
    for kernel in kernels:
        t0 = time.perf_counter()
        plan = ops.tile_plan(conn0, xy, pat, kernel=kernel)
        torch.cuda.synchronize()
        out[kernel + "_plan_s"] = time.perf_counter() - t0
        if plan is None:
            out[kernel] = None
            continue
        out[kernel + "_stats"] = {k: (v if not isinstance(v, float) else round(v, 4)) for k, v in plan.stats.items()}
        v = torch.full((pat.nnz,), float("nan"), dtype=torch.float64, device="cuda")
        ops.assemble_values(conn0, xy, pat, plan, v)
        torch.cuda.synchronize()
        eq = bool(torch.equal(v, v_rows))
        if not eq:
            bad = (v != v_rows) | torch.isnan(v)
            out[kernel + "_bad"] = int(bad.sum().item())
            out[kernel + "_nan"] = int(torch.isnan(v).sum().item())
            idx = torch.nonzero(bad).flatten()[:8].tolist()
            out[kernel + "_first_bad"] = idx
        for _ in range(3):
            ops.assemble_values(conn0, xy, pat, plan, v)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            ops.assemble_values(conn0, xy, pat, plan, v)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        out[kernel] = {"ms": ms, "equal_rows": eq, "gbs": bytes_alg / ms / 1e6, "frac": bytes_alg / ms / 1e6 / 6454.6}
        del v, plan
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
