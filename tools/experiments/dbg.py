import os, sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from feprogram_b200 import mesh, ops
n = 4096
conn_h, coords_h, bc = mesh.structured_quad4(n, dtype=np.int32)
conn = torch.from_numpy(conn_h - 1).cuda(); coords = torch.from_numpy(coords_h).cuda()
pat = ops.symbolic(conn, coords_h.shape[0], 4, 2)
plan = ops.tile_plan(conn, coords, pat)
print(plan.stats)
vals = torch.empty(pat.nnz, dtype=torch.float64, device='cuda')
def timeit(label, f, reps=5):
    for _ in range(2): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): f()
    b.record(); torch.cuda.synchronize()
    print('%-40s %.3f ms' % (label, a.elapsed_time(b) / reps))
for dbg in ('0', '1', '2'):
    os.environ['FE_TILED_DBG'] = dbg
    timeit('tiled dbg=' + dbg, lambda: ops.assemble_values(conn, coords, pat, plan, vals))
os.environ['FE_TILED_DBG'] = '0'
timeit('rows', lambda: ops.assemble_values(conn, coords, pat, None, vals))
plan2 = ops.tile_plan(conn, coords, pat, morton=False)
print(plan2.stats if plan2 else None)
if plan2: timeit("tiled identity order", lambda: ops.assemble_values(conn, coords, pat, plan2, vals))
