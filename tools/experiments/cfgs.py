import json, os, sys, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
RESULTS = []
from feprogram_b200 import mesh, ops
def run(name, kind, n, shuffle=False, perturb=False, op=0, renumber=False):
    gen = {4: mesh.structured_quad4, 9: mesh.structured_quad9, 3: mesh.structured_tri3}[kind]
    conn_h, coords_h, bc = gen(n, dtype=np.int32)
    mx = 2 * n if kind == 9 else n
    if perturb: coords_h = mesh.perturb_interior(coords_h, mx, mx, 0.2, 0)
    if shuffle: conn_h, coords_h, bc, _ = mesh.shuffle_numbering(conn_h, coords_h, bc, 1)
    conn = torch.from_numpy((conn_h - 1).astype(np.int32)).cuda(); coords = torch.from_numpy(coords_h).cuda()
    nnode = coords_h.shape[0]
    if renumber:
        torch.cuda.synchronize(); a = time.perf_counter()
        perm, order = ops.locality_order(coords)
        conn = perm[conn.to(torch.int64)].to(torch.int32).contiguous(); coords = coords[order].contiguous()
        torch.cuda.synchronize(); print('renumbering %.1f ms' % (1e3 * (time.perf_counter() - a)))
    pat = ops.symbolic(conn, nnode, kind, 2); plan = ops.tile_plan(conn, coords, pat); torch.cuda.synchronize()
    a = time.perf_counter(); pat = ops.symbolic(conn, nnode, kind, 2); plan = ops.tile_plan(conn, coords, pat); torch.cuda.synchronize()
    sym = time.perf_counter() - a
    vals = torch.empty(pat.nnz, dtype=torch.float64, device='cuda')
    def t(f, reps=5):
        for _ in range(2): f()
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): f()
        e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / reps
    ms = t(lambda: ops.assemble_values(conn, coords, pat, plan, vals, op))
    ab = conn.numel() * 4 + nnode * 16 + pat.nnz * 8
    x = torch.randn(pat.nrows, dtype=torch.float64, device='cuda'); y = torch.empty_like(x)
    sm = t(lambda: torch.ops.feprogram.spmv(pat.row_ptr, pat.col_idx, vals, x, y, None), 10)
    sb = 12 * pat.nnz + 20 * pat.nrows
    print('%-28s elems %.1fM nnz %.0fM | symbolic %.1f ms | assemble[%s] %.3f ms = %.2f Gelem/s, %.0f GB/s (%.2f of 6454.6) | spmv %.3f ms %.0f GB/s (%.2f) | plan %s'
          % (name, conn.shape[0] / 1e6, pat.nnz / 1e6, 1e3 * sym, 'tiled' if plan else 'rows', ms, conn.shape[0] / ms / 1e6, ab / ms / 1e6, ab / ms / 1e6 / 6454.6,
             sm, sb / sm / 1e6, sb / sm / 1e6 / 6454.6, None if plan is None else {k: plan.stats[k] for k in ('max_nodes', 'max_halo', 'max_work', 'cell_scale', 'mode', 'ntiles')}), flush=True)
    RESULTS.append({"config": name, "elem_type": kind, "operator": ["reference", "weighted", "mass"][op], "elements": int(conn.shape[0]),
                    "nnz": int(pat.nnz), "symbolic_ms": 1e3 * sym, "assembly_kernel": "k_assemble_tiled" if plan else "k_assemble_rows",
                    "assembly_ms": ms, "elements_per_s": conn.shape[0] / ms * 1e3, "assembly_gbs": ab / ms / 1e6,
                    "assembly_frac_of_hbm_peak": ab / ms / 1e6 / 6454.6, "spmv_ms": sm, "spmv_gbs": sb / sm / 1e6,
                    "spmv_frac_of_hbm_peak": sb / sm / 1e6 / 6454.6, "shuffled": shuffle, "perturbed": perturb, "solver_renumbering": renumber})
    del vals, pat, plan, conn, coords, x, y; torch.cuda.empty_cache()
which = sys.argv[1:] or ['q4', 'q9', 't3', 't3s', 'q4p']
if 'q4' in which: run('cfg2 Quad4 4096^2', 4, 4096)
if 'q4p' in which: run('Quad4 4096^2 perturbed', 4, 4096, perturb=True)
if 'q9' in which: run('cfg3 Quad9 2048^2', 9, 2048)
if 't3' in which: run('Tri3 3500^2x2', 3, 3500)
if 't3s' in which: run('Tri3 3500^2x2 shuffled+pert', 3, 3500, shuffle=True, perturb=True)
if 'q4big' in which: run('cfg4 Quad4 8192^2 (1 GPU)', 4, 8192)
if 'q9m' in which: run('cfg3 Quad9 2048^2 mass', 9, 2048, op=2)
if 'q4m' in which: run('cfg2 Quad4 4096^2 mass', 4, 4096, op=2)
if 'cfg5' in which: run('cfg5 Tri3 5000^2x2 shuffled+perturbed (1 GPU)', 3, 5000, shuffle=True, perturb=True)
if 'cfg5r' in which: run('cfg5 Tri3 5000^2x2 shuffled+perturbed, Morton-renumbered inside the solver (1 GPU)', 3, 5000, shuffle=True, perturb=True, renumber=True)
os.makedirs('gpurun_out', exist_ok=True)
with open('gpurun_out/cfgs.json', 'w') as f:
    json.dump(RESULTS, f, indent=1)
