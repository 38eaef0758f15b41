"""torchrun worker: where does a multi-GPU e2e step (DistFemProblem from pinned host arrays) spend its time?"""
import os, sys, time, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from feprogram_b200 import dist as fdist, ops
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
n = 4096
m, bc = fdist.structured_quad4_block(n, n, rank, world)
comm = fdist.DistFemProblem.make_comm(rank, world, dev)
conn_p = torch.from_numpy(m.conn).pin_memory(); xy_p = torch.from_numpy(m.coords).pin_memory()
host = fdist.LocalMesh(**{**m.__dict__, "conn": conn_p.numpy(), "coords": xy_p.numpy()})
def T():
    torch.cuda.synchronize(); return time.perf_counter()
for it in range(3):
    t0 = T()
    conn = torch.from_numpy(np.ascontiguousarray(host.conn)).to(dev); coords = torch.from_numpy(host.coords).to(dev)
    t1 = T()
    pat = ops.symbolic(conn, host.nnode, 4, 2)
    t2 = T()
    plan = ops.tile_plan(conn, coords, pat)
    t3 = T()
    q = fdist.DistFemProblem(host, dev, 4, comm)
    t4 = T()
    q.set_bc(bc["dir"], bc["neu"]); q.assemble()
    t5 = T()
    if rank == 0:
        print("it %d: h2d %.1f  symbolic %.1f  plan %.1f (%s, %d tiles)  ctor(all) %.1f  bc+assemble %.1f ms" % (
            it, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), plan.stats.get("mode") if plan else None, plan.ntiles if plan else 0,
            1e3 * (t4 - t3), 1e3 * (t5 - t4)), flush=True)
    del conn, coords, pat, plan, q
dist.barrier(); dist.destroy_process_group()
