#!/usr/bin/env python3
"""Per-warp event time stamps of k_assemble_ws (CTA 0, first 512 tiles) from a -DFE_WS_TRACE build -- developer tool.

    FE_LIB=tools/experiments/variants/lib_trace.so python tools/experiments/trace_ws.py [KIND] [N] [OP]

Writes gpurun_out/ws_trace_<kind>_<op>.npy: int64 [16 warps][512 tiles][8 events] (SM clock)."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    import torch
    from feprogram_b200 import _lib
    _lib.LIB_PATH = os.path.join(ROOT, os.environ["FE_LIB"])
    from feprogram_b200 import mesh, ops
    kind = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    op = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    gen = {4: mesh.structured_quad4, 3: mesh.structured_tri3}[kind]
    conn, coords, _ = gen(n, dtype=np.int32)
    conn0 = torch.from_numpy(conn - 1).cuda()
    xy = torch.from_numpy(coords).cuda()
    pat = ops.symbolic(conn0, coords.shape[0], kind, 1 if op == 3 else 2)
    plan = ops.tile_plan(conn0, xy, pat, kernel="ws")
    v = torch.empty(pat.nnz, dtype=torch.float64, device="cuda")
    for _ in range(4):
        ops.assemble_values(conn0, xy, pat, plan, v, op)
    torch.cuda.synchronize()
    out = np.zeros((16, 512, 8), dtype=np.int64)
    L = _lib.lib()
    L.fe_ws_trace_read.argtypes = [C.c_void_p, C.c_size_t]
    _lib.check(L.fe_ws_trace_read(out.ctypes.data_as(C.c_void_p), out.nbytes))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    np.save(os.path.join(ROOT, "gpurun_out", "ws_trace_%d_%d.npy" % (kind, op)), out)
    print("saved", out.shape, int(out.max() - out[out > 0].min()))


if __name__ == "__main__":
    main()
