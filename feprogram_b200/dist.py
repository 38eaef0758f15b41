"""Multi-GPU scale-out of the FEM hot path: one process per GPU, NCCL over NVLink (north_star).

The reference has no distributed code.  Partition = contiguous element blocks in the reference's element
order.  Rank r OWNS the nodes whose first incident element lies in its block and assembles exactly those
rows from its own elements plus the recomputed ghost elements that touch them -- assembly needs no
communication and every owned row is summed in the same (ascending global element id) order as on one
GPU, so the values are bit-identical to the single-GPU matrix.  The solve exchanges the search direction
on interface nodes once per iteration and all-reduces two scalars (csrc/fe_dist.cuh).

Local numbering of rank r: owned nodes first (ascending global id), then ghost nodes grouped by owner
(ascending rank, ascending global id) so each neighbour's values land in one contiguous vector segment.

Host-side partition logic is NumPy (one-time set-up, also exercised by the CPU gloo tests); everything
that touches values runs in the CUDA library.
"""
from __future__ import annotations

import ctypes as C
import os
import time
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

__all__ = ["LocalMesh", "element_blocks", "partition_mesh", "structured_quad4_block", "exchange_send_lists",
           "DistFemProblem"]


@dataclass
class LocalMesh:
    """One rank's share of a mesh (host arrays)."""
    rank: int
    world: int
    conn: np.ndarray            # (nelem_local, nnpe) int32, LOCAL 0-based node ids, ascending global element id
    elem_gid: np.ndarray        # (nelem_local,) int64
    n_own_elems: int            # elements of this rank's own block (the rest are ghost elements)
    coords: np.ndarray          # (nnode_local, 2) fp64
    node_gid: np.ndarray        # (nnode_local,) int64: owned then ghosts grouped by owner (ascending gid, or Z-curve)
    n_owned: int
    ghost_owner: np.ndarray     # (n_ghost,) int32 owner rank of each ghost node (non-decreasing)
    # filled by exchange_send_lists(): neighbours and the node-level send lists
    nbr: List[int] = field(default_factory=list)
    send_nodes: List[np.ndarray] = field(default_factory=list)   # per neighbour: local owned node ids
    recv_range: List[tuple] = field(default_factory=list)        # per neighbour: (ghost_start, ghost_end) local node ids

    @property
    def nnode(self) -> int:
        return self.coords.shape[0]

    def halo_arrays(self, ndof: int = 2):
        """dof-level halo description for fe_dist_*: nbr_rank, send_off, send_idx, recv_off."""
        nbr = np.asarray(self.nbr, dtype=np.int32)
        send_idx, send_off = [], [0]
        recv_off = []
        for k in range(len(self.nbr)):
            nodes = np.asarray(self.send_nodes[k], dtype=np.int64)
            dofs = (nodes[:, None] * ndof + np.arange(ndof)[None, :]).ravel()
            send_idx.append(dofs)
            send_off.append(send_off[-1] + dofs.size)
        # ghost segments are laid out in neighbour order; neighbours without ghosts get an empty segment
        pos = self.n_owned * ndof
        for k in range(len(self.nbr)):
            g0, g1 = self.recv_range[k]
            if g1 > g0:
                pos = g0 * ndof
            recv_off.append(pos)
            pos += (g1 - g0) * ndof
        recv_off.append(pos)
        si = np.concatenate(send_idx).astype(np.int32) if send_idx else np.zeros(0, np.int32)
        return nbr, np.asarray(send_off, np.int64), si, np.asarray(recv_off, np.int64)


def element_blocks(nelem: int, world: int) -> np.ndarray:
    """Boundaries of the contiguous element blocks: block r = [b[r], b[r+1])."""
    return (np.arange(world + 1, dtype=np.int64) * nelem) // world


def morton_keys(xy: np.ndarray, lo: np.ndarray, span: np.ndarray) -> np.ndarray:
    """Z-curve keys (16 bits per axis) of points inside the box (lo, lo + span)."""
    q = np.clip(((xy - lo) / np.maximum(span, 1e-300) * 65535.0).astype(np.int64), 0, 65535)

    def spread(v):
        v = (v | (v << 8)) & 0x00FF00FF
        v = (v | (v << 4)) & 0x0F0F0F0F
        v = (v | (v << 2)) & 0x33333333
        v = (v | (v << 1)) & 0x55555555
        return v

    return spread(q[:, 0]) | (spread(q[:, 1]) << 1)


def partition_mesh(conn0: np.ndarray, coords: np.ndarray, rank: int, world: int,
                   bounds: Optional[np.ndarray] = None, locality: bool = False) -> LocalMesh:
    """General (unstructured, any numbering) partition from the full mesh.  conn0 is GLOBAL 0-based.

    ``locality``: number the rank's owned nodes (and each owner's group of ghosts) along a Z-curve instead of
    by ascending global id -- for meshes whose global numbering has no locality (BASELINE cfg5), so that the
    local CSR rows of an assembly tile are contiguous and the SpMV's gathers stay cached."""
    conn0 = np.asarray(conn0, dtype=np.int64)
    nelem, nnpe = conn0.shape
    nnode = coords.shape[0]
    bounds = element_blocks(nelem, world) if bounds is None else np.asarray(bounds, dtype=np.int64)
    first = np.full(nnode, nelem, dtype=np.int64)
    # lowest incident element of every node: element-major writes in DESCENDING element order, the last
    # (= lowest) one stays (what np.minimum.at computes, two orders of magnitude faster on 10^8 entries)
    first[conn0[::-1].ravel()] = np.repeat(np.arange(nelem - 1, -1, -1, dtype=np.int64), nnpe)
    owner = np.searchsorted(bounds, first, side="right") - 1
    owner[first == nelem] = -1                                   # nodes without elements belong to nobody
    mine = owner == rank
    need = mine[conn0].any(axis=1)                                # every element touching an owned node
    egid = np.nonzero(need)[0]
    n_own = int(((egid >= bounds[rank]) & (egid < bounds[rank + 1])).sum())
    touched = np.unique(conn0[egid])
    ghosts = touched[owner[touched] != rank]
    ghosts = ghosts[np.lexsort((ghosts, owner[ghosts]))]
    owned = np.nonzero(mine)[0]
    if locality:
        lo = coords[touched, :2].min(axis=0)
        span = coords[touched, :2].max(axis=0) - lo
        owned = owned[np.argsort(morton_keys(coords[owned, :2], lo, span), kind="stable")]
        ghosts = ghosts[np.lexsort((morton_keys(coords[ghosts, :2], lo, span), owner[ghosts]))]
    node_gid = np.concatenate([owned, ghosts]).astype(np.int64)
    g2l = np.full(nnode, -1, dtype=np.int64)
    g2l[node_gid] = np.arange(node_gid.size)
    return LocalMesh(rank, world, g2l[conn0[egid]].astype(np.int32), egid.astype(np.int64), n_own,
                     np.ascontiguousarray(coords[node_gid, :2], dtype=np.float64), node_gid, int(owned.size),
                     owner[ghosts].astype(np.int32))


def structured_quad4_block(nx: int, ny_per_rank: int, rank: int, world: int) -> tuple:
    """Rank's share of the structured nx x (world*ny_per_rank) Quad4 mesh on [0,1]^2, built locally
    (no global arrays): returns (LocalMesh with halo lists, bc dict with local Dirichlet / load dofs)."""
    ny = world * ny_per_rank
    w = nx + 1
    j0, j1 = rank * ny_per_rank, (rank + 1) * ny_per_rank          # own element rows [j0, j1)
    last = rank == world - 1
    own_rows = np.arange(0 if rank == 0 else j0 + 1, j1 + 1)       # owned node rows
    ghost_lo = np.array([j0]) if rank > 0 else np.zeros(0, np.int64)
    ghost_hi = np.array([j1 + 1]) if not last else np.zeros(0, np.int64)
    node_rows = np.concatenate([own_rows, ghost_lo, ghost_hi]).astype(np.int64)
    i = np.arange(w, dtype=np.int64)
    node_gid = (node_rows[:, None] * w + i[None, :]).ravel()
    row_pos = {int(r): k for k, r in enumerate(node_rows)}
    xs = np.linspace(0.0, 1.0, nx + 1)
    ys = np.linspace(0.0, 1.0, ny + 1)
    coords = np.empty((node_gid.size, 2))
    coords[:, 0] = np.tile(xs, node_rows.size)
    coords[:, 1] = np.repeat(ys[node_rows], w)
    erows = np.arange(j0, j1 + (0 if last else 1), dtype=np.int64)  # own rows + the ghost row above
    ie = np.arange(nx, dtype=np.int64)
    lo = np.array([row_pos[int(j)] for j in erows], dtype=np.int64)[:, None] * w + ie[None, :]
    hi = np.array([row_pos[int(j) + 1] for j in erows], dtype=np.int64)[:, None] * w + ie[None, :]
    conn = np.stack([lo, lo + 1, hi + 1, hi], axis=2).reshape(-1, 4).astype(np.int32)
    egid = (erows[:, None] * nx + ie[None, :]).ravel()
    n_owned = own_rows.size * w
    ghost_owner = np.concatenate([np.full(ghost_lo.size * w, rank - 1), np.full(ghost_hi.size * w, rank + 1)]).astype(np.int32)
    m = LocalMesh(rank, world, conn, egid, ny_per_rank * nx, coords, node_gid, n_owned, ghost_owner)
    # halo lists: the row below my first owned row is rank-1's ghost "hi" row, etc.
    if rank > 0:
        m.nbr.append(rank - 1)
        m.send_nodes.append(row_pos[j0 + 1] * w + i)               # they need my first owned row
        m.recv_range.append((n_owned, n_owned + w))
    if not last:
        m.nbr.append(rank + 1)
        m.send_nodes.append(row_pos[j1] * w + i)                   # they need my last owned row
        g0 = n_owned + ghost_lo.size * w
        m.recv_range.append((g0, g0 + w))
    # boundary conditions of the reference's set-up: bottom edge Dirichlet, left edge nodal load (both dofs)
    bottom = (row_pos[0] * w + i) if rank == 0 else np.zeros(0, np.int64)
    left = np.array([row_pos[int(r)] * w for r in own_rows], dtype=np.int64)
    dofs = lambda nodes: np.sort(np.concatenate([2 * nodes, 2 * nodes + 1])).astype(np.int64)  # noqa: E731
    return m, {"dir": dofs(bottom), "neu": dofs(left)}


def exchange_send_lists(m: LocalMesh, group=None) -> LocalMesh:
    """Completes a LocalMesh from partition_mesh with neighbour / send / receive lists.  Collective over
    torch.distributed (any backend): every rank tells each owner which of its nodes it needs."""
    import torch.distributed as dist
    need = {}
    n_ghost = m.nnode - m.n_owned
    for q in np.unique(m.ghost_owner).tolist():
        sel = np.nonzero(m.ghost_owner == q)[0]
        need[int(q)] = (m.node_gid[m.n_owned + sel].tolist(), int(m.n_owned + sel[0]), int(m.n_owned + sel[-1] + 1))
    gathered = [None] * m.world
    dist.all_gather_object(gathered, {q: v[0] for q, v in need.items()}, group=group)
    sends = {q: gathered[q][m.rank] for q in range(m.world) if q != m.rank and m.rank in gathered[q]}
    owned_gid = m.node_gid[: m.n_owned]
    sorter = np.argsort(owned_gid, kind="stable")              # owned nodes need not be in ascending gid order
    sorted_gid = owned_gid[sorter]
    m.nbr, m.send_nodes, m.recv_range = [], [], []
    for q in sorted(set(need) | set(sends)):
        m.nbr.append(q)
        gids = np.asarray(sends.get(q, []), dtype=np.int64)
        pos = np.minimum(np.searchsorted(sorted_gid, gids), max(0, sorted_gid.size - 1))
        loc = sorter[pos] if gids.size else np.zeros(0, dtype=np.int64)
        assert np.array_equal(owned_gid[loc], gids), "a neighbour asked for a node this rank does not own"
        m.send_nodes.append(loc)
        m.recv_range.append(need[q][1:] if q in need else (m.n_owned + n_ghost, m.n_owned + n_ghost))
    return m


def peer_ghost_tables(m: LocalMesh, group=None):
    """For every ghost node of this rank: (owner rank, node index inside the owner) -- what a peer-memory SpMV
    needs to read the owner's copy directly.  Collective over torch.distributed (any backend): every owner
    publishes what it sends to whom (owner-local node indices, in the requester's ghost order)."""
    import torch.distributed as dist
    lists = [None] * m.world
    dist.all_gather_object(lists, {int(q): np.asarray(s_).tolist() for q, s_ in zip(m.nbr, m.send_nodes)}, group=group)
    n_ghost = m.nnode - m.n_owned
    grank = np.zeros(max(1, n_ghost), dtype=np.int32)
    gridx = np.zeros(max(1, n_ghost), dtype=np.int32)
    for q, (a, b) in zip(m.nbr, m.recv_range):
        if b > a:
            theirs = np.asarray(lists[q][m.rank], dtype=np.int32)
            assert theirs.size == b - a
            grank[a - m.n_owned: b - m.n_owned] = q
            gridx[a - m.n_owned: b - m.n_owned] = theirs
    return grank, gridx


# ----------------------------------------------------------------------------------------------------
# device side
# ----------------------------------------------------------------------------------------------------
class DistFemProblem:
    """Assembly of the owned rows + distributed Jacobi-PCG for one rank (CUDA + NCCL only)."""

    def __init__(self, local: LocalMesh, device, elem_type: int = 4, comm=None, timing: bool = True, bc=None,
                 rhs_out=None):
        """``timing``: bracket the symbolic phase with host synchronisations so that ``symbolic_s`` is its wall time
        on resident connectivity (the bench's resident-input line); False = no synchronisation beyond the phase's own
        read-backs (``symbolic_s`` is then None), as FemProblem.assemble does.
        ``bc = (dir_dofs_local, neu_dofs_local[, value])``: boundary dof sets known up front -- the right-hand side and
        the Dirichlet mask do not depend on K, so they are built on a side stream while the mesh is still uploading
        and, with ``rhs_out`` (a pinned host tensor of 2 * n_owned doubles), the owned right-hand side travels to the
        host at once (what FemProblem.assemble does on one GPU); without ``bc`` call ``set_bc`` afterwards."""
        import torch
        from . import ops
        from .elements import side_streams
        self.m = local
        self.dev = torch.device(device)
        self.elem_type = elem_type
        self.comm = comm
        torch.cuda.set_device(self.dev)
        self.nrows_owned = 2 * local.n_owned
        self.nrows_local = 2 * local.nnode
        main = torch.cuda.current_stream(self.dev)
        bc_event = None
        if bc is not None:                               # tiny uploads, ahead of the mesh on the copy engine
            d = torch.from_numpy(np.asarray(bc[0], dtype=np.int64)).to(self.dev)
            n = torch.from_numpy(np.asarray(bc[1], dtype=np.int64)).to(self.dev)
            value = float(bc[2]) if len(bc) > 2 else 1.0
            bcs = side_streams(self.dev)[1]
            bcs.wait_event(main.record_event())
            with torch.cuda.stream(bcs):
                self.rhs = torch.empty(self.nrows_local, dtype=torch.float64, device=self.dev)
                self.dirmask = torch.empty(self.nrows_local, dtype=torch.uint8, device=self.dev)
                torch.ops.feprogram.bc_rhs(d, n, value, self.rhs, self.dirmask)
                if rhs_out is not None:
                    rhs_out.copy_(self.rhs[: self.nrows_owned], non_blocking=True)
                bc_event = bcs.record_event()
            for t in (self.rhs, self.dirmask):
                t.record_stream(main)
            for t in (d, n):
                t.record_stream(bcs)
            self._bc = (d, n, value)
        # connectivity first; the coordinates follow on a side stream while the CSR pattern (which only needs the
        # connectivity) is built -- as FemProblem does (elements.py, _state)
        self.conn = torch.from_numpy(np.ascontiguousarray(local.conn)).to(self.dev, non_blocking=True)
        side = side_streams(self.dev)[0]
        side.wait_event(main.record_event())
        with torch.cuda.stream(side):
            self.coords = torch.from_numpy(np.ascontiguousarray(local.coords)).to(self.dev, non_blocking=True)
        self.coords.record_stream(main)
        coords_ready = side.record_event()
        if timing:
            main.synchronize()                           # connectivity has arrived; coordinates still in flight
        t0 = time.perf_counter()
        # a lattice numbering's tiles are formed on a side stream while the first symbolic kernels run (ops.tile_formation)
        csr_side = side_streams(self.dev)[2]
        formed = {}

        def form_tiles(lattice):
            formed["tiles"] = ops.tile_formation(self.conn.shape[0], local.nnode, elem_type, lattice, csr_side)

        def early_layout(eptr, einc, nptr):          # first half of the plan, concurrent with fe_adjacency_fill
            formed["layout"] = ops.plan_layout_early(self.conn, elem_type, local.nnode, eptr, einc, nptr,
                                                     formed.get("tiles"), csr_side)
        # (the dof-level row_ptr / col_idx are expanded when the SpMV / PCG calls first read them: Pattern.row_ptr)
        self.pat = ops.symbolic(self.conn, local.nnode, elem_type, 2, csr="lazy", between=form_tiles,
                                after_fill=early_layout)
        self.plan = ops.tile_plan(self.conn, self.coords, self.pat, coords_ready=lambda: main.wait_event(coords_ready),
                                  formation=formed.get("tiles"), layout=formed.get("layout"))
        main.wait_event(coords_ready)
        self.pat.csr_ready()
        self.symbolic_s = None
        if timing:
            torch.cuda.synchronize(self.dev)
            self.symbolic_s = time.perf_counter() - t0  # may include the tail of the coordinate upload
        self.vals = torch.empty(self.pat.nnz, dtype=torch.float64, device=self.dev)
        self.nnz_owned = 4 * int(self.pat.nptr[local.n_owned].item())   # host copy: the SpMV / PCG calls take it as an argument
        nbr, send_off, send_idx, recv_off = local.halo_arrays(2)
        self._nbr, self._send_off, self._recv_off = nbr, send_off, recv_off       # host arrays (kept alive)
        self.send_idx = torch.from_numpy(send_idx).to(self.dev)
        self.sendbuf = torch.empty(max(1, int(send_off[-1])), dtype=torch.float64, device=self.dev)
        if bc_event is None:
            self.rhs = torch.zeros(self.nrows_local, dtype=torch.float64, device=self.dev)
            self.dirmask = torch.zeros(self.nrows_local, dtype=torch.uint8, device=self.dev)
        else:
            main.wait_event(bc_event)

    # -- communicator ------------------------------------------------------------------------------
    @staticmethod
    def make_comm(rank: int, world: int, device):
        """ncclComm_t for the library (its own communicator; the id travels over torch.distributed)."""
        import torch
        import torch.distributed as dist
        from . import _lib
        L = _lib.lib()
        if not L.fe_dist_available():
            raise RuntimeError("libfeprogram_b200.so was built without NCCL")
        uid = (C.c_uint8 * 128)()
        if rank == 0:
            _lib.check(L.fe_dist_unique_id(uid))
        t = torch.tensor(list(uid), dtype=torch.uint8, device=device if dist.get_backend() == "nccl" else "cpu")
        dist.broadcast(t, 0)
        uid = (C.c_uint8 * 128)(*t.cpu().tolist())
        comm = C.c_void_p()
        with torch.cuda.device(device):
            _lib.check(L.fe_dist_comm_init(uid, rank, world, C.byref(comm)))
        return comm

    # -- assembly -----------------------------------------------------------------------------------
    def assemble(self):
        from . import ops
        ops.assemble_values(self.conn, self.coords, self.pat, self.plan, self.vals, 0)

    def set_bc(self, dir_dofs_local: np.ndarray, neu_dofs_local: np.ndarray, value: float = 1.0):
        import torch
        d = torch.from_numpy(np.asarray(dir_dofs_local, dtype=np.int64)).to(self.dev)
        n = torch.from_numpy(np.asarray(neu_dofs_local, dtype=np.int64)).to(self.dev)
        torch.ops.feprogram.bc_rhs(d, n, float(value), self.rhs, self.dirmask)
        self._bc = (d, n, float(value))

    def apply_bc(self):
        """The per-step BC work (rhs + Dirichlet mask), as FemProblem.assemble does after the numeric pass."""
        import torch
        d, n, v = self._bc
        torch.ops.feprogram.bc_rhs(d, n, v, self.rhs, self.dirmask)

    # -- halo / solve ----------------------------------------------------------------------------------
    def _halo_args(self):
        p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        return (len(self._nbr), p(self._nbr), p(self._send_off), C.c_void_p(self.send_idx.data_ptr()),
                C.c_void_p(self.sendbuf.data_ptr()), p(self._recv_off))

    def halo_exchange(self, vec):
        import torch
        from . import _lib
        n, nbr, soff, sidx, sbuf, roff = self._halo_args()
        with torch.cuda.device(self.dev):
            st = C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
            _lib.check(_lib.lib().fe_dist_halo_exchange(self.comm, n, nbr, soff, sidx, sbuf, roff,
                                                        C.c_void_p(vec.data_ptr()), st))

    def spmv_owned(self, x, y):
        """y[owned rows] = A x with x valid on owned + ghost dofs (call halo_exchange(x) first)."""
        import torch
        from . import _lib
        with torch.cuda.device(self.dev):
            st = C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
            pat = self.pat
            _lib.check(_lib.lib().fe_spmv(self.nrows_owned, self.nnz_owned, int(pat.row_ptr.dtype == torch.int64),
                                          C.c_void_p(pat.row_ptr.data_ptr()), C.c_void_p(pat.col_idx.data_ptr()),
                                          C.c_void_p(self.vals.data_ptr()), C.c_void_p(x.data_ptr()),
                                          C.c_void_p(y.data_ptr()), None, st))

    def enable_p2p(self, group=None):
        """Collective.  Sets up the peer-memory halo for solve(): a double-buffered direction vector and a flag
        array per rank in CUDA symmetric memory (peer-mapped over NVLink), plus, for every ghost node, its
        owner and its index inside the owner."""
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        m = self.m
        group = group or dist.group.WORLD
        # The peer-memory kernels wait inside a kernel for flags other ranks write: that is only live when every
        # rank runs on its OWN GPU and all of them can map each other's memory.  Check both collectively and
        # stay on the NCCL path otherwise (returns False; solve() then uses fe_dist_pcg).
        props = torch.cuda.get_device_properties(self.dev)
        mine = (str(getattr(props, "uuid", "")) or "%s/%d" % (os.uname().nodename, self.dev.index), self.dev.index)
        devs = [None] * m.world
        dist.all_gather_object(devs, mine, group=group)
        distinct = len({d[0] for d in devs}) == m.world
        peer_ok = all(q == self.dev.index or torch.cuda.can_device_access_peer(self.dev.index, q)
                      for _, q in devs) if distinct else False
        oks = [None] * m.world
        dist.all_gather_object(oks, bool(distinct and peer_ok), group=group)
        if not all(oks):
            self.p2p = False
            return False
        sizes = [None] * m.world
        dist.all_gather_object(sizes, self.nrows_local, group=group)
        self._pstride = (max(sizes) + 15) // 16 * 16
        self._pbuf = symm_mem.empty(2 * self._pstride, dtype=torch.float64, device=self.dev)
        self._flags = symm_mem.empty(max(m.world, 8), dtype=torch.int64, device=self.dev)
        self._red = symm_mem.empty(64, dtype=torch.float64, device=self.dev)      # double[2][16][2] slot table
        self._rver = symm_mem.empty(16, dtype=torch.int64, device=self.dev)
        for t in (self._pbuf, self._flags, self._red, self._rver):
            t.zero_()
        hp = symm_mem.rendezvous(self._pbuf, group)
        hf = symm_mem.rendezvous(self._flags, group)
        hr = symm_mem.rendezvous(self._red, group)
        hv = symm_mem.rendezvous(self._rver, group)
        self._symm_handles = (hp, hf, hr, hv)
        self._peer_red = torch.tensor(list(hr.buffer_ptrs), dtype=torch.int64, device=self.dev)
        self._peer_rver = torch.tensor(list(hv.buffer_ptrs), dtype=torch.int64, device=self.dev)
        self._seq = 0
        self._peer_p = torch.tensor(list(hp.buffer_ptrs), dtype=torch.int64, device=self.dev)
        self._peer_flags = torch.tensor(list(hf.buffer_ptrs), dtype=torch.int64, device=self.dev)
        grank, gridx = peer_ghost_tables(m, group)
        self._ghost_rank = torch.from_numpy(grank).to(self.dev)
        self._ghost_ridx = torch.from_numpy(gridx).to(self.dev)
        self._nbr_dev = torch.from_numpy(np.ascontiguousarray(self._nbr, dtype=np.int32)).to(self.dev) \
            if len(self._nbr) else torch.zeros(1, dtype=torch.int32, device=self.dev)
        self._epoch = 0
        torch.cuda.synchronize(self.dev)
        dist.barrier(group=group)
        self.p2p = True
        return True

    def solve(self, rtol: float = 1e-10, maxit: int = 100000, check_every: int = 50, p2p=None):
        import torch
        from . import _lib
        L = _lib.lib()
        pat = self.pat
        if p2p is None:
            p2p = getattr(self, "p2p", False)
        if p2p:
            return self._solve_p2p(rtol, maxit, check_every)
        x = torch.zeros(self.nrows_local, dtype=torch.float64, device=self.dev)
        work = torch.empty(L.fe_dist_pcg_workspace_bytes(self.nrows_local), dtype=torch.uint8, device=self.dev)
        info = np.zeros(4)
        n, nbr, soff, sidx, sbuf, roff = self._halo_args()
        with torch.cuda.device(self.dev):
            st = C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
            _lib.check(L.fe_dist_pcg(self.nrows_owned, self.nrows_local, self.nnz_owned, int(pat.row_ptr.dtype == torch.int64),
                                     C.c_void_p(pat.row_ptr.data_ptr()), C.c_void_p(pat.col_idx.data_ptr()),
                                     C.c_void_p(self.vals.data_ptr()), C.c_void_p(self.rhs.data_ptr()),
                                     C.c_void_p(self.dirmask.data_ptr()), C.c_void_p(x.data_ptr()), float(rtol), 0.0,
                                     int(maxit), int(check_every), C.c_void_p(work.data_ptr()),
                                     info.ctypes.data_as(C.c_void_p), self.comm, n, nbr, soff, sidx, sbuf, roff,
                                     C.c_void_p(pat.nptr.data_ptr()), C.c_void_p(pat.nadj.data_ptr()), st))
        self.info = dict(iterations=int(info[0]), relres=float(info[1]), r0=float(info[2]), converged=info[3] == 1.0,
                         breakdown=info[3] < 0.0)
        return x


    def _solve_p2p(self, rtol, maxit, check_every):
        import torch
        from . import _lib
        L = _lib.lib()
        pat = self.pat
        x = torch.zeros(self.nrows_local, dtype=torch.float64, device=self.dev)
        work = torch.empty(L.fe_dist_pcg_workspace_bytes(self.nrows_local), dtype=torch.uint8, device=self.dev)
        info = np.zeros(4)
        n, nbr, soff, sidx, sbuf, roff = self._halo_args()
        dp = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
        with torch.cuda.device(self.dev):
            st = C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
            epoch, seq = self._epoch, self._seq
            self._epoch += int(maxit) + 2
            self._seq += 2 * int(maxit) + 2
            _lib.check(L.fe_dist_pcg_p2p(self.nrows_owned, self.nrows_local, int(pat.row_ptr.dtype == torch.int64),
                                         dp(pat.row_ptr), dp(pat.col_idx), dp(self.vals), dp(self.rhs), dp(self.dirmask),
                                         dp(x), float(rtol), 0.0, int(maxit), int(check_every), dp(work),
                                         info.ctypes.data_as(C.c_void_p), self.comm, n, nbr, soff, sidx, sbuf, roff,
                                         dp(pat.nptr), dp(pat.nadj), self.m.rank, self.m.world, dp(self._peer_p),
                                         dp(self._peer_flags), dp(self._ghost_rank), dp(self._ghost_ridx),
                                         dp(self._nbr_dev), self._pstride, dp(self._pbuf), epoch,
                                         dp(self._peer_red), dp(self._peer_rver), seq, st))
        self.info = dict(iterations=int(info[0]), relres=float(info[1]), r0=float(info[2]), converged=info[3] == 1.0,
                         breakdown=info[3] < 0.0)
        return x


# ----------------------------------------------------------------------------------------------------
# bench.py support: weak-scaling workload, rank r = one n x n element block of the n x (world*n) mesh
# ----------------------------------------------------------------------------------------------------
class _BenchPart:
    def __init__(self, n: int, rank: int, world: int, dev, ny_per_rank: int, halo: str = "p2p", comm=None):
        import torch
        self.torch = torch
        self.m, self.bc = structured_quad4_block(n, ny_per_rank, rank, world)
        self.comm = comm if comm is not None else DistFemProblem.make_comm(rank, world, dev)
        warm, _ = structured_quad4_block(64, 64, rank, world)     # tiny problem first: CUDA module load, first mallocs
        DistFemProblem(warm, dev, 4, self.comm).assemble()
        self.p = DistFemProblem(self.m, dev, 4, self.comm)
        self.p.set_bc(self.bc["dir"], self.bc["neu"])
        # one-time symbolic phase of this rank's mesh, timed on resident inputs like the single-GPU line of bench.py:
        # median of three rebuilds of pattern + tile plan (the first build above also paid the first cudaMallocs)
        from . import ops
        times = []
        for _ in range(3):
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            pat = ops.symbolic(self.p.conn, self.m.nnode, 4, 2)
            plan = ops.tile_plan(self.p.conn, self.p.coords, pat)
            torch.cuda.synchronize(dev)
            times.append(time.perf_counter() - t0)
            del pat, plan
        self.symbolic_s = sorted(times)[1]
        self.halo = "nccl"
        if world > 1 and halo == "p2p" and self.p.enable_p2p():
            self.halo = "peer-memory (NVLink loads + flag-gated scalar exchange)"
        self.nelem_owned = self.m.n_own_elems
        self.tiled = self.p.plan is not None
        self.launches_per_step = 3
        pat = self.p.pat
        nnz_owned = 4 * int(pat.nptr[self.m.n_owned].item())
        self.asm_bytes = self.m.conn.shape[0] * 16 + self.m.nnode * 16 + pat.nnz * 8
        self.spmv_bytes = 12 * nnz_owned + 20 * self.p.nrows_owned
        self.dev = dev

    def assemble_step(self):
        self.p.assemble()
        self.p.apply_bc()

    def assemble_kernel_only(self):
        self.p.assemble()

    def _sync(self):
        import torch.distributed as dist
        dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def time_spmv(self, warm: int, reps: int) -> float:
        torch = self.torch
        x = torch.randn(self.p.nrows_local, dtype=torch.float64, device=self.dev)
        y = torch.empty_like(x)
        for _ in range(warm):
            self.p.halo_exchange(x)
            self.p.spmv_owned(x, y)
        self._sync()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            self.p.halo_exchange(x)
            self.p.spmv_owned(x, y)
        b.record()
        self._sync()
        return a.elapsed_time(b) / reps

    def time_pcg(self, iters: int):
        torch = self.torch
        self.p.solve(rtol=0.0, maxit=10, check_every=10)
        self._sync()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        self.p.solve(rtol=0.0, maxit=iters, check_every=iters)
        b.record()
        self._sync()
        return a.elapsed_time(b), self.p.info["iterations"]

    def e2e(self, K: int):
        """Through the public multi-GPU API with HOST inputs: upload of the rank's connectivity and
        coordinates (pinned), symbolic phase, numeric assembly, BCs, download of the owned rhs + checksum."""
        import torch.distributed as dist
        torch = self.torch
        conn_p = torch.from_numpy(self.m.conn).pin_memory()
        xy_p = torch.from_numpy(self.m.coords).pin_memory()
        host = LocalMesh(**{**self.m.__dict__, "conn": conn_p.numpy(), "coords": xy_p.numpy()})
        out = torch.empty(self.p.nrows_owned, dtype=torch.float64).pin_memory()

        def step():
            # the right-hand side does not depend on K: built and sent to the host while the mesh is still uploading
            q = DistFemProblem(host, self.dev, 4, self.comm, timing=False, bc=(self.bc["dir"], self.bc["neu"]), rhs_out=out)
            q.assemble()
            return q.vals.sum().item()
        step()
        self._sync()
        reps = max(1, min(K, 3))
        t0 = time.perf_counter()
        for _ in range(reps):
            chk = step()
        self._sync()
        dt = torch.tensor([(time.perf_counter() - t0) / reps], dtype=torch.float64, device=self.dev)
        tot = torch.tensor([float(self.nelem_owned)], dtype=torch.float64, device=self.dev)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        return {"value": float(tot.item() / dt.item()), "unit": "elements/s",
                "h2d_bytes_per_step": int(self.m.conn.nbytes + self.m.coords.nbytes + self.bc["dir"].nbytes + self.bc["neu"].nbytes),
                "d2h_bytes_per_step": int(self.p.nrows_owned * 8 + 8), "ms_per_step": 1e3 * float(dt.item()),
                "includes": "per rank: H2D, symbolic phase, numeric assembly, BCs, D2H rhs + checksum", "checksum": chk}


def build_structured_block(n: int, rank: int, world: int, dev, ny_per_rank: int = 0, halo: str = "p2p", comm=None):
    return _BenchPart(n, rank, world, dev, ny_per_rank or n, halo, comm)
