"""GPU side of the multi-GPU path.  (1) On one GPU: every rank's share assembled in turn -- owned rows are
bit-identical to the single-domain matrix.  (2) With >= 2 GPUs: fe_dist_pcg under torchrun (2 ranks, NCCL)
against the single-GPU solve."""
import os
import subprocess
import sys

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import ROOT
from oracle import fe_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world,kind", [(2, 4), (3, 4), (4, 4), (2, 9)])
def test_rank_shares_reproduce_single_domain_rows_bitwise(world, kind):
    import torch
    from feprogram_b200 import build, mesh
    build.build()
    from feprogram_b200 import dist as fdist, ops
    gen = mesh.structured_quad4 if kind == 4 else mesh.structured_quad9
    n = 40 if kind == 4 else 12
    conn, coords, bc = gen(n)
    mx = n if kind == 4 else 2 * n
    coords = mesh.perturb_interior(coords, mx, mx, 0.2, seed=5)
    conn, coords, bc, _ = mesh.shuffle_numbering(conn, coords, bc, seed=6)
    conn0 = torch.from_numpy((conn - 1).astype(np.int32)).cuda()
    xy = torch.from_numpy(coords).cuda()
    pat = ops.symbolic(conn0, coords.shape[0], kind, 2)
    vals = torch.empty(pat.nnz, dtype=torch.float64, device="cuda")
    ops.assemble_values(conn0, xy, pat, ops.tile_plan(conn0, xy, pat), vals)
    K = sp.csr_matrix((vals.cpu().numpy(), pat.col_idx.cpu().numpy(), pat.row_ptr.cpu().numpy()),
                      shape=(pat.nrows, pat.nrows))
    seen = np.zeros(coords.shape[0], dtype=int)
    for rank in range(world):
        m = fdist.partition_mesh(conn - 1, coords, rank, world, locality=(world == 3 or kind == 9))
        p = fdist.DistFemProblem(m, "cuda:0", kind, comm=None)
        p.assemble()
        Kl = sp.csr_matrix((p.vals.cpu().numpy(), p.pat.col_idx.cpu().numpy(), p.pat.row_ptr.cpu().numpy()),
                           shape=(p.nrows_local, p.nrows_local))
        gdof = (2 * m.node_gid[:, None] + np.arange(2)[None, :]).ravel()
        seen[m.node_gid[: m.n_owned]] += 1
        for r in range(p.nrows_owned):
            cl = gdof[Kl.indices[Kl.indptr[r]:Kl.indptr[r + 1]]]
            o = np.argsort(cl)
            gr = gdof[r]
            assert np.array_equal(cl[o], K.indices[K.indptr[gr]:K.indptr[gr + 1]])
            assert np.array_equal(Kl.data[Kl.indptr[r]:Kl.indptr[r + 1]][o], K.data[K.indptr[gr]:K.indptr[gr + 1]])
    assert np.all(seen == 1)


WORKER = r"""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %(root)r)
from feprogram_b200 import mesh, dist as fdist
from feprogram_b200.elements import FemProblem
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
for case in ("general", "structured"):
    if case == "general":
        conn, coords, bc = mesh.structured_quad4(48)
        coords = mesh.perturb_interior(coords, 48, 48, 0.2, seed=2)
        conn, coords, bc, _ = mesh.shuffle_numbering(conn, coords, bc, seed=3)
        m = fdist.exchange_send_lists(fdist.partition_mesh(conn - 1, coords, rank, world, locality=True))
        d = np.array(sorted(bc[0]), dtype=np.int64) - 1; l = np.array(sorted(bc[3]), dtype=np.int64) - 1
        g2l = np.full(coords.shape[0], -1, dtype=np.int64); g2l[m.node_gid] = np.arange(m.nnode)
        loc = lambda nodes: np.sort(np.concatenate([2 * g2l[nodes][g2l[nodes] >= 0], 2 * g2l[nodes][g2l[nodes] >= 0] + 1]))
        bcl = {"dir": loc(d), "neu": loc(l)}
    else:
        conn, coords, bc = mesh.structured_quad4(32, 32 * world)
        m, bcl = fdist.structured_quad4_block(32, 32, rank, world)
    comm = fdist.DistFemProblem.make_comm(rank, world, dev)
    p = fdist.DistFemProblem(m, dev, 4, comm)
    p.set_bc(bcl["dir"], bcl["neu"])
    p.assemble()
    x = p.solve(rtol=1e-11)
    x2 = p.solve(rtol=1e-11)
    assert torch.equal(x, x2), "distributed solve is not reproducible"
    it_nccl = p.info["iterations"]
    p.enable_p2p()                      # ghost entries read from the owner's memory over NVLink
    for _ in range(2):
        x3 = p.solve(rtol=1e-11)
        assert p.info["iterations"] == it_nccl and torch.equal(x, x3), "peer-memory halo changed the iterates"
    x = p.solve(rtol=1e-11, p2p=False)
    # single-GPU reference on every rank
    fem = FemProblem(conn, coords, device=dev); fem.setMatrix(); fem.setBoundaryConditions(bc, verbose=False)
    fem.assemble(); fem.rtol = 1e-11; U = fem.solveProblem()
    gdof = (2 * m.node_gid[:, None] + np.arange(2)[None, :]).ravel()
    err = np.linalg.norm(x.cpu().numpy() - U[gdof]) / np.linalg.norm(U)
    print(case, "rank", rank, "iters", p.info["iterations"], "single", fem.solver_info["iterations"], "err", err, flush=True)
    assert p.info["converged"] and err < 1e-8, err
dist.barrier()
dist.destroy_process_group()
print("DIST_OK", rank, flush=True)
"""


def test_distributed_pcg_two_ranks_nccl(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    env = dict(os.environ)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and r.stdout.count("DIST_OK") == 2, r.stdout[-3000:] + r.stderr[-3000:]
