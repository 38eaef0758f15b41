"""Host-side logic of the multi-GPU path on CPU: world_size 2 and 3 over gloo.  Each rank partitions the
mesh, exchanges halo lists, builds its local matrix with the oracle, and runs the distributed Jacobi-PCG
algorithm (halo exchange of p + two scalar all-reduces per iteration) in NumPy; the result must match the
global direct solve, and the owned rows must be the global matrix's rows."""
import os
import socket

import numpy as np
import pytest
import scipy.sparse as sp
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT  # noqa: F401
from feprogram_b200 import dist as fdist
from feprogram_b200 import mesh
from oracle import fe_oracle as orc


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _halo(m, halo, vec):
    nbr, send_off, send_idx, recv_off = halo
    reqs, keep = [], []
    for k, q in enumerate(nbr.tolist()):
        sbuf = torch.from_numpy(vec[send_idx[send_off[k]:send_off[k + 1]]].copy())
        keep.append(sbuf)
        if sbuf.numel():
            reqs.append(dist.isend(sbuf, q))
    recvs = []
    for k, q in enumerate(nbr.tolist()):
        n = int(recv_off[k + 1] - recv_off[k])
        if n:
            rbuf = torch.empty(n, dtype=torch.float64)
            recvs.append((k, rbuf))
            reqs.append(dist.irecv(rbuf, q))
    for r in reqs:
        r.wait()
    for k, rbuf in recvs:
        vec[recv_off[k]:recv_off[k + 1]] = rbuf.numpy()


def _allsum(*vals):
    t = torch.tensor(vals, dtype=torch.float64)
    dist.all_reduce(t)
    return t.tolist()


def _worker(rank, world, port, kind, structured, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        if structured:
            nx, nyr = 7, 3
            conn, coords, bc = mesh.structured_quad4(nx, world * nyr)
            m, bcl = fdist.structured_quad4_block(nx, nyr, rank, world)
            ref = fdist.exchange_send_lists(fdist.partition_mesh(conn - 1, coords, rank, world))
            assert np.array_equal(m.node_gid, ref.node_gid) and m.n_owned == ref.n_owned
            assert np.array_equal(m.elem_gid, ref.elem_gid) and np.array_equal(m.conn, ref.conn)
            assert np.allclose(m.coords, ref.coords) and m.nbr == ref.nbr
            for a, b in zip(m.send_nodes, ref.send_nodes):
                assert np.array_equal(np.asarray(a), np.asarray(b))
            assert [tuple(r) for r in m.recv_range] == [tuple(r) for r in ref.recv_range]
        else:
            gen = mesh.structured_quad4 if kind == 4 else mesh.structured_quad9
            n = 9 if kind == 4 else 4
            conn, coords, bc = gen(n)
            mx = n if kind == 4 else 2 * n
            coords = mesh.perturb_interior(coords, mx, mx, 0.2, seed=3)
            conn, coords, bc, _ = mesh.shuffle_numbering(conn, coords, bc, seed=8)
            # Quad9 case: Z-curve local numbering (owned nodes no longer in ascending global id)
            m = fdist.exchange_send_lists(fdist.partition_mesh(conn - 1, coords, rank, world, locality=(kind == 9)))
            if kind == 9 and m.n_owned > 2:
                assert not np.all(np.diff(m.node_gid[: m.n_owned]) > 0)
        conn0 = conn - 1
        K = orc.assemble_csr(conn0, coords)
        d, nn = orc.boundary_dofs(bc)
        Kbc, b = orc.apply_bc_reference(K, d, nn)
        U = orc.solve_direct(Kbc, b)
        # every node is owned exactly once
        owned_all = [None] * world
        dist.all_gather_object(owned_all, m.node_gid[: m.n_owned].tolist())
        allo = np.sort(np.concatenate([np.asarray(o, dtype=np.int64) for o in owned_all]))
        assert np.array_equal(allo, np.arange(coords.shape[0]))
        # owned rows of the local matrix == rows of the global matrix
        Kl = orc.assemble_csr(m.conn, m.coords).tocsr()
        no = 2 * m.n_owned
        gdof = (2 * m.node_gid[:, None] + np.arange(2)[None, :]).ravel()
        Kg = K.tocsr()
        for r in range(0, no, max(1, no // 40)):
            cl = gdof[Kl.indices[Kl.indptr[r]:Kl.indptr[r + 1]]]
            vl = Kl.data[Kl.indptr[r]:Kl.indptr[r + 1]]
            o = np.argsort(cl)
            gr = gdof[r]
            assert np.array_equal(cl[o], Kg.indices[Kg.indptr[gr]:Kg.indptr[gr + 1]])
            assert np.allclose(vl[o], Kg.data[Kg.indptr[gr]:Kg.indptr[gr + 1]], rtol=0, atol=1e-13 * np.abs(K.data).max())
        # distributed Jacobi-PCG (the algorithm of csrc/fe_dist.cuh) in NumPy over gloo
        halo = m.halo_arrays(2)
        nl = 2 * m.nnode
        bl = b[gdof]
        isdir = np.zeros(K.shape[0], bool)
        isdir[d] = True
        mask = isdir[gdof]
        A = Kl[:no]
        x = np.zeros(nl)
        x[:no] = np.where(mask[:no], bl[:no], 0.0)
        _halo(m, halo, x)
        r_ = np.where(mask[:no], 0.0, bl[:no] - A @ x)
        diag = Kl.diagonal()[:no]
        minv = np.where(mask[:no] | (diag == 0), 0.0, 1.0 / np.where(diag == 0, 1.0, diag))
        p = np.zeros(nl)
        p[:no] = minv * r_
        rz, rr0 = _allsum(float(r_ @ (minv * r_)), float(r_ @ r_))
        rr, it = rr0, 0
        while rr > 1e-24 * rr0 and it < 5000:
            _halo(m, halo, p)
            Ap = A @ p
            Ap[mask[:no]] = 0.0
            (pAp,) = _allsum(float(p[:no] @ Ap))
            alpha = rz / pAp
            x[:no] += alpha * p[:no]
            r_ -= alpha * Ap
            z = minv * r_
            rz_new, rr = _allsum(float(r_ @ z), float(r_ @ r_))
            p[:no] = z + (rz_new / rz) * p[:no]
            rz = rz_new
            it += 1
        _halo(m, halo, x)
        err = np.linalg.norm(x - U[gdof]) / np.linalg.norm(U)
        assert err < 1e-8, err
        # peer-memory tables (fe_dist_pcg_p2p): ghost g lives at node gridx[g] of rank grank[g]
        grank, gridx = fdist.peer_ghost_tables(m)
        gids = [None] * world
        dist.all_gather_object(gids, m.node_gid[: m.n_owned].tolist())
        ng = m.nnode - m.n_owned
        assert np.array_equal(grank[:ng], m.ghost_owner)
        for g in range(ng):
            assert gids[grank[g]][gridx[g]] == m.node_gid[m.n_owned + g]
        out.put((rank, "ok", it))
    except Exception as e:  # noqa: BLE001
        import traceback
        out.put((rank, "fail", traceback.format_exc() + repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,kind,structured", [(2, 4, False), (3, 4, False), (2, 9, False), (3, 4, True)])
def test_partition_halo_and_distributed_pcg_over_gloo(world, kind, structured):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, kind, structured, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    bad = [r for r in res if r[1] != "ok"]
    assert not bad, bad[0][2]
    assert len({r[2] for r in res}) == 1          # every rank ran the same number of iterations


def test_row_block_local_numbering_fits_the_generalised_lattice_rule():
    """fe_lattice_keys (csrc/fe_tile_form.cu) lets a rank's tile plan start before its coordinates arrive when every
    element is {b, b+1, c+1, c} with b, c in the same column of two node rows of W consecutive ids (W read from the
    middle element; strict row-major form iff the first element has c = b + W).  This restates the rule on the host
    and checks that the local meshes structured_quad4_block builds satisfy it with one distinct cell per element."""
    from feprogram_b200 import dist as fd

    def lattice_cells(conn):
        c0, cm = conn[0], conn[len(conn) // 2]
        W = int(cm[3] - cm[0])
        strict = int(c0[3] - c0[0]) == W
        x, y, z, w = conn.T.astype(np.int64)
        ok = (W >= 2) & (y == x + 1) & (z == w + 1) & (x >= 0) & (w >= 0)
        rx, rw = x // W, w // W
        ix = x - rx * W
        ok &= (ix == w - rw * W) & (rx != rw) & ((not strict) | (rw == rx + 1)) & (ix < W - 1)
        return bool(ok.all()), W, strict, ix, np.where(strict, rx, rw)

    for world in (1, 2, 3, 5):
        for rank in range(world):
            m, _ = fd.structured_quad4_block(24, 10, rank, world)
            ok, W, strict, ix, iy = lattice_cells(m.conn)
            assert ok and W == 25 and strict == (rank == 0)
            assert len(set(zip(ix.tolist(), iy.tolist()))) == len(m.conn)          # one cell per element
            assert iy.min() >= 0 and iy.max() < (m.nnode + W - 1) // W             # inside the key range the host sizes
    # a shuffled numbering does not fit
    from feprogram_b200 import mesh
    conn, coords, bc = mesh.structured_quad4(12, 9)
    conn2, _, _, _ = mesh.shuffle_numbering(conn, coords, bc, seed=4)
    assert not lattice_cells(conn2 - 1)[0]
