import os, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
t = symm_mem.empty(1024, dtype=torch.float64, device=dev)
t.fill_(float(rank))
hdl = symm_mem.rendezvous(t, dist.group.WORLD)
print(rank, "buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs], "signal", [hex(p) for p in hdl.signal_pad_ptrs][:2], flush=True)
dist.barrier()
peer = (rank + 1) % world
pb = hdl.get_buffer(peer, (1024,), torch.float64)
print(rank, "peer value read over NVLink:", float(pb[3].item()), flush=True)
dist.barrier()
pb[rank] = 100.0 + rank
torch.cuda.synchronize(); dist.barrier()
print(rank, "my buffer after peer write:", t[:world].tolist(), flush=True)
dist.destroy_process_group()
