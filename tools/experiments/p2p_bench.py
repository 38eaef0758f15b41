"""torchrun worker: NCCL-halo vs peer-memory-halo PCG on the weak-scaling block (n x n elements per rank)."""
import os, sys, time, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from feprogram_b200 import dist as fdist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
nyr = int(sys.argv[2]) if len(sys.argv) > 2 else n
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 200
m, bc = fdist.structured_quad4_block(n, nyr, rank, world)
comm = fdist.DistFemProblem.make_comm(rank, world, dev)
p = fdist.DistFemProblem(m, dev, 4, comm)
p.set_bc(bc["dir"], bc["neu"]); p.assemble()
def run(p2p):
    p.solve(rtol=0.0, maxit=10, check_every=10, p2p=p2p)
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); x = p.solve(rtol=0.0, maxit=iters, check_every=iters, p2p=p2p); b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b)], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return x, float(t.item())
x0, t0 = run(False)
p.enable_p2p()
x1, t1 = run(True)
x2, t2 = run(True)
if rank == 0:
    print("n", n, "x", nyr, "world", world, "iters", iters, "nccl ms/it %.4f  p2p ms/it %.4f %.4f  equal %s" % (t0 / iters, t1 / iters, t2 / iters, bool(torch.equal(x0, x1))), flush=True)
dist.barrier(); dist.destroy_process_group()
