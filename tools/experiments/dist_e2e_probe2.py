"""torchrun worker: the bench's own multi-GPU e2e step, timed piece by piece after the bench's setup."""
import os, sys, time, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from feprogram_b200 import dist as fdist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
part = fdist.build_structured_block(4096, rank, world, dev, 4096)
part.assemble_step(); part.time_spmv(3, 5); part.time_pcg(20)
r = part.e2e(3)
if rank == 0: print("bench e2e ms_per_step %.1f" % r["ms_per_step"], flush=True)
conn_p = torch.from_numpy(part.m.conn).pin_memory(); xy_p = torch.from_numpy(part.m.coords).pin_memory()
host = fdist.LocalMesh(**{**part.m.__dict__, "conn": conn_p.numpy(), "coords": xy_p.numpy()})
out = torch.empty(part.p.nrows_owned, dtype=torch.float64).pin_memory()
def T():
    torch.cuda.synchronize(); return time.perf_counter()
for it in range(3):
    t0 = T(); q = fdist.DistFemProblem(host, dev, 4, part.comm)
    t1 = T(); q.set_bc(part.bc["dir"], part.bc["neu"]); q.assemble()
    t2 = T(); out.copy_(q.rhs[: q.nrows_owned], non_blocking=True)
    t3 = T(); c = q.vals.sum().item()
    t4 = T(); del q
    t5 = T()
    if rank == 0:
        print("it %d: ctor %.1f  bc+assemble %.1f  d2h %.1f  checksum %.1f  free %.1f ms" % (it, *(1e3 * (b - a) for a, b in ((t0, t1), (t1, t2), (t2, t3), (t3, t4), (t4, t5)))), flush=True)
dist.barrier(); dist.destroy_process_group()
