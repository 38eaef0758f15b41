"""torchrun worker for BASELINE cfg5: perturbed Tri3 mesh, ~50M elements, randomly shuffled node numbering,
element-block partition over the ranks (general partition_mesh path, Z-curve local numbering), owned-row
assembly + peer-memory PCG.  Rank 0 writes one JSON object to gpurun_out/cfg5_dist.json."""
import json, os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from feprogram_b200 import mesh, dist as fdist

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
locality = (sys.argv[2] != "0") if len(sys.argv) > 2 else True
iters = 200
t0 = time.time()
conn, coords, bc = mesh.structured_tri3(n, dtype=np.int32)
coords = mesh.perturb_interior(coords, n, n, 0.2, 0)
conn, coords, bc, _ = mesh.shuffle_numbering(conn, coords, bc, 1)
t1 = time.time()
m = fdist.exchange_send_lists(fdist.partition_mesh(conn - 1, coords, rank, world, locality=locality))
t2 = time.time()
g2l = np.full(coords.shape[0], -1, dtype=np.int64)
g2l[m.node_gid] = np.arange(m.nnode)
def loc(tags):
    l = g2l[np.fromiter((int(t) for t in tags), dtype=np.int64) - 1]
    l = l[l >= 0]
    return np.sort(np.concatenate([2 * l, 2 * l + 1]))
nelem_total, nnode_total = conn.shape[0], coords.shape[0]
dir_l, neu_l = loc(bc[0]), loc(bc[3])
del conn, coords, g2l
comm = fdist.DistFemProblem.make_comm(rank, world, dev)
p = fdist.DistFemProblem(m, dev, 3, comm)
p = fdist.DistFemProblem(m, dev, 3, comm)
p.set_bc(dir_l, neu_l)
p.enable_p2p()
def timed(f, reps):
    for _ in range(3): f()
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): f()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / reps], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
ms_asm = timed(p.assemble, 5)
x = torch.randn(p.nrows_local, dtype=torch.float64, device=dev); y = torch.empty_like(x)
ms_spmv = timed(lambda: p.spmv_owned(x, y), 10)
p.solve(rtol=0.0, maxit=10, check_every=10)
ms_pcg = timed(lambda: p.solve(rtol=0.0, maxit=iters, check_every=iters), 1)
stats = torch.tensor([p.pat.nnz, m.conn.shape[0], m.n_owned, m.nnode - m.n_owned], dtype=torch.float64, device=dev)
allst = [torch.empty_like(stats) for _ in range(world)]
dist.all_gather(allst, stats)
if rank == 0:
    nnz = sum(float(s[0]) for s in allst)
    nrows = 2 * nnode_total
    asm_bytes = nelem_total * 3 * 4 + nnode_total * 16 + nnz * 8
    spmv_bytes = 12 * nnz + 20 * nrows
    out = {"config": "cfg5 Tri3 %dx%dx2 perturbed, shuffled node numbering, dp%d element blocks" % (n, n, world),
           "n_gpus": world, "elements": nelem_total, "nnz_owned_total": nnz, "local_numbering": "z-curve" if locality else "ascending global id",
           "assembly_ms": ms_asm, "elements_per_s": nelem_total / ms_asm * 1e3,
           "assembly_gbs_per_gpu": asm_bytes / world / ms_asm / 1e6, "assembly_frac_of_hbm_peak": asm_bytes / world / ms_asm / 1e6 / 6454.6,
           "spmv_ms": ms_spmv, "spmv_gbs_per_gpu": spmv_bytes / world / ms_spmv / 1e6, "spmv_frac_of_hbm_peak": spmv_bytes / world / ms_spmv / 1e6 / 6454.6,
           "pcg_iters_per_s": iters / ms_pcg * 1e3, "halo": "peer-memory", "symbolic_phase_s": p.symbolic_s,
           "ghost_nodes_per_rank": [int(s[3]) for s in allst], "ghost_elements_per_rank": [int(s[1]) - nelem_total // world for s in allst],
           "host_generate_s": t1 - t0, "host_partition_s": t2 - t1, "tile_plan": None if p.plan is None else p.plan.stats.get("mode")}
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/cfg5_dist.json", "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out), flush=True)
dist.barrier(); dist.destroy_process_group()
