import os, sys, numpy as np, torch, subprocess
sys.path.insert(0, '/root/repo')
if len(sys.argv) > 1:
    from feprogram_b200 import mesh, ops
    n = 4096
    conn_h, coords_h, bc = mesh.structured_quad4(n, dtype=np.int32)
    conn = torch.from_numpy(conn_h - 1).cuda(); coords = torch.from_numpy(coords_h).cuda()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    pat = ops.symbolic(conn, coords_h.shape[0], 4, 2); plan = ops.tile_plan(conn, coords, pat)
    torch.cuda.synchronize()
    import time
    a = time.perf_counter(); pat = ops.symbolic(conn, coords_h.shape[0], 4, 2); torch.cuda.synchronize(); b = time.perf_counter()
    plan = ops.tile_plan(conn, coords, pat); torch.cuda.synchronize(); c = time.perf_counter()
    print('symbolic %.1f ms, tile plan %.1f ms' % (1e3 * (b - a), 1e3 * (c - b)))
    vals = torch.empty(pat.nnz, dtype=torch.float64, device='cuda')
    ops.assemble_values(conn, coords, pat, plan, vals)
    x = torch.randn(pat.nrows, dtype=torch.float64, device='cuda'); y = torch.empty_like(x)
    for _ in range(3): torch.ops.feprogram.spmv(pat.row_ptr, pat.col_idx, vals, x, y, None)
    torch.cuda.synchronize(); t0.record()
    for _ in range(10): torch.ops.feprogram.spmv(pat.row_ptr, pat.col_idx, vals, x, y, None)
    t1.record(); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / 10
    print('LPR', os.environ.get('FE_SPMV_LPR'), 'spmv %.3f ms %.0f GB/s' % (ms, (12 * pat.nnz + 20 * pat.nrows) / ms / 1e6))
else:
    for lpr in ('2', '4', '8', '16', '32'):
        env = dict(os.environ, FE_SPMV_LPR=lpr)
        print(subprocess.run([sys.executable, __file__, 'x'], env=env, capture_output=True, text=True).stdout.strip())
