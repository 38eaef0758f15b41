import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from feprogram_b200 import mesh, ops
for n, shuffle in ((40, False), (6, False)):
    conn, coords, bc = mesh.structured_tri3(n, n + 3)
    conn0 = torch.from_numpy((conn - 1).astype(np.int32)).cuda(); xy = torch.from_numpy(coords).cuda()
    pat = ops.symbolic(conn0, coords.shape[0], 3, 2)
    v_rows = torch.empty(pat.nnz, dtype=torch.float64, device="cuda")
    ops.assemble_values(conn0, xy, pat, None, v_rows)
    for morton in (True, False):
        plan = ops.tile_plan(conn0, xy, pat, morton=morton)
        print(n, morton, None if plan is None else plan.stats)
        if plan is None: continue
        v = torch.full((pat.nnz,), float("nan"), dtype=torch.float64, device="cuda")
        ops.assemble_values(conn0, xy, pat, plan, v)
        bad = (v != v_rows) | v.isnan()
        print('  mismatches', int(bad.sum()), 'of', pat.nnz, 'nan', int(v.isnan().sum()), 'maxdiff', float((v - v_rows).abs().nan_to_num().max()))
        if bad.any():
            idx = bad.nonzero()[:10, 0].cpu().numpy()
            rp = pat.row_ptr.cpu().numpy()
            rows = np.searchsorted(rp, idx, side='right') - 1
            print('  first bad slots', idx, 'rows', rows, 'cols', pat.col_idx.cpu().numpy()[idx])
            print('  got', v[idx].cpu().numpy(), 'want', v_rows[idx].cpu().numpy())
