import os, sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from feprogram_b200 import mesh, ops
n = int(os.environ.get("N", "4096"))
conn_h, coords_h, bc = mesh.structured_quad4(n, dtype=np.int32)
conn = torch.from_numpy(conn_h - 1).cuda(); coords = torch.from_numpy(coords_h).cuda()
pat = ops.symbolic(conn, coords_h.shape[0], 4, 2)
plan = ops.tile_plan(conn, coords, pat)
vals = torch.empty(pat.nnz, dtype=torch.float64, device='cuda')
for _ in range(4):
    ops.assemble_values(conn, coords, pat, plan, vals)
torch.cuda.synchronize()
