"""PyTorch custom ops over the C ABI (include/feprogram_b200.h).

PyTorch is plumbing here: device memory, streams and (for multi-GPU) torch.distributed.  Every op
launches hand-written sm_100a kernels from libfeprogram_b200.so on torch's current CUDA stream;
nothing falls back to torch or the CPU.

    torch.ops.feprogram.element_matrices / symbolic pieces / assemble / bc_rhs / apply_bc /
    spmv / csr_diagonal / pcg

plus thin Python helpers (``symbolic``, ``Pattern``) that chain the one-time symbolic steps.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import (FE_BC_ROWS, FE_BC_SYMMETRIC, FE_OP_MASS, FE_OP_POISSON, FE_OP_REFERENCE, FE_OP_WEIGHTED,  # noqa: F401
                   FE_QUAD4, FE_QUAD9, FE_TRI3)

Tensor = torch.Tensor


def _ptr(t: Optional[Tensor]):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _stream(t: Tensor):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _chk_dev(*ts: Tensor):
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("feprogram_b200 ops need CUDA tensors (no CPU fallback); got device %s" % t.device)
        if not t.is_contiguous():
            raise RuntimeError("feprogram_b200 ops need contiguous tensors")


def _read_small(parts, dtype) -> list:
    t = parts[0] if len(parts) == 1 else torch.cat([p.reshape(-1) for p in parts])
    t = t.contiguous()
    if t.dtype != dtype:
        raise TypeError("expected %s, got %s" % (dtype, t.dtype))
    host = torch.empty(t.numel(), dtype=dtype, pin_memory=True)
    with torch.cuda.device(t.device):
        _lib.check(_lib.lib().fe_peek(_ptr(t), C.c_void_p(host.data_ptr()), t.element_size() * t.numel(), _stream(t)))
        torch.cuda.current_stream(t.device).synchronize()
    return host.tolist()


def read_i32(*parts: Tensor) -> list:
    """Host copy of a few int32 device values (sizes, status words): ONE synchronisation, no copy engine (fe_peek stores
    into pinned host memory from a kernel, so the read does not wait behind a large download on another stream)."""
    return _read_small(parts, torch.int32)


def read_f64(*parts: Tensor) -> list:
    """Same for a few fp64 device values."""
    return _read_small(parts, torch.float64)


def _want(t: Tensor, dtype, name: str):
    if t.dtype != dtype:
        raise TypeError("%s must be %s, got %s" % (name, dtype, t.dtype))


# --------------------------------------------------------------------------------------------
# host: tables
# --------------------------------------------------------------------------------------------
def tables(elem_type: int):
    """(dH[ng2,2,nnpe], H[ng2,nnpe], w[ng2]) from the library's host-side fe_tables."""
    nnpe, ng2 = C.c_int32(0), C.c_int32(0)
    _lib.check(_lib.lib().fe_tables(elem_type, None, None, None, C.byref(nnpe), C.byref(ng2)))
    dH = np.empty((ng2.value, 2, nnpe.value))
    H = np.empty((ng2.value, nnpe.value))
    w = np.empty(ng2.value)
    _lib.check(_lib.lib().fe_tables(elem_type, dH.ctypes.data_as(C.c_void_p), H.ctypes.data_as(C.c_void_p),
                                    w.ctypes.data_as(C.c_void_p), None, None))
    return dH, H, w


# --------------------------------------------------------------------------------------------
# custom ops
# --------------------------------------------------------------------------------------------
@torch.library.custom_op("feprogram::tags_to_conn", mutates_args=())
def tags_to_conn(tags: Tensor) -> Tensor:
    """1-based integer tags (int32 / int64 / uint64 bit pattern) -> int32 0-based connectivity."""
    _chk_dev(tags)
    if tags.element_size() not in (4, 8) or tags.dtype.is_floating_point:
        raise TypeError("tags must be 4- or 8-byte integers, got %s" % tags.dtype)
    with torch.cuda.device(tags.device):
        conn = torch.empty(tags.shape, dtype=torch.int32, device=tags.device)
        _lib.check(_lib.lib().fe_tags_to_conn(tags.element_size(), tags.numel(), _ptr(tags), _ptr(conn),
                                              _stream(tags)))
    return conn


@torch.library.custom_op("feprogram::element_matrices", mutates_args=())
def element_matrices(conn: Tensor, coords: Tensor, elem_type: int, op: int) -> Tensor:
    """Dense Ke per element, (nelem, ndof*nnpe, ndof*nnpe) fp64 (ndof = 1 for FE_OP_POISSON, else 2).  fe_element_matrices."""
    _chk_dev(conn, coords)
    _want(conn, torch.int32, "conn")
    _want(coords, torch.float64, "coords")
    nelem, nnpe = conn.shape
    m = nnpe if op == FE_OP_POISSON else 2 * nnpe
    with torch.cuda.device(conn.device):
        ke = torch.empty((nelem, m, m), dtype=torch.float64, device=conn.device)
        _lib.check(_lib.lib().fe_element_matrices(elem_type, op, nelem, _ptr(conn), _ptr(coords), _ptr(ke),
                                                  _stream(conn)))
    return ke


@torch.library.custom_op("feprogram::gauss_gradients", mutates_args=())
def gauss_gradients(conn: Tensor, coords: Tensor, U: Tensor, elem_type: int) -> Tensor:
    """(nelem, ng2, 4) fp64: (du/dx, du/dy, dv/dx, dv/dy) at every Gauss point.  fe_gauss_gradients."""
    _chk_dev(conn, coords, U)
    _want(conn, torch.int32, "conn")
    _want(coords, torch.float64, "coords")
    _want(U, torch.float64, "U")
    nelem, nnpe = conn.shape
    ng2 = 1 if nnpe == 3 else nnpe
    if U.numel() != 2 * coords.shape[0]:
        raise ValueError("U must hold 2 values per node")
    with torch.cuda.device(conn.device):
        out = torch.empty((nelem, ng2, 4), dtype=torch.float64, device=conn.device)
        _lib.check(_lib.lib().fe_gauss_gradients(elem_type, nelem, _ptr(conn), _ptr(coords), _ptr(U), _ptr(out),
                                                 _stream(conn)))
    return out


@torch.library.custom_op("feprogram::nodal_strain", mutates_args=())
def nodal_strain(conn: Tensor, coords: Tensor, U: Tensor, eptr: Tensor, einc: Tensor, elem_type: int) -> Tensor:
    """(nnode, 3) fp64 nodal means of (du/dx, dv/dy, du/dy + dv/dx).  fe_nodal_strain."""
    _chk_dev(conn, coords, U, eptr, einc)
    _want(conn, torch.int32, "conn")
    _want(coords, torch.float64, "coords")
    _want(U, torch.float64, "U")
    nnode = coords.shape[0]
    if U.numel() != 2 * nnode:
        raise ValueError("U must hold 2 values per node")
    with torch.cuda.device(conn.device):
        out = torch.empty((nnode, 3), dtype=torch.float64, device=conn.device)
        _lib.check(_lib.lib().fe_nodal_strain(elem_type, nnode, _ptr(eptr), _ptr(einc), _ptr(conn), _ptr(coords),
                                              _ptr(U), _ptr(out), _stream(conn)))
    return out


@torch.library.custom_op("feprogram::incidence", mutates_args=())
def incidence(conn: Tensor, nnode: int, elem_type: int) -> Tuple[Tensor, Tensor]:
    """node -> (element*16 + local node) incidence: (eptr[nnode+1], einc[nelem*nnpe]) int32."""
    _chk_dev(conn)
    _want(conn, torch.int32, "conn")
    nelem, nnpe = conn.shape
    dev = conn.device
    with torch.cuda.device(dev):
        L = _lib.lib()
        eptr = torch.empty(nnode + 1, dtype=torch.int32, device=dev)
        ws = torch.empty(max(1, L.fe_scan_workspace_bytes(nnode + 1)), dtype=torch.uint8, device=dev)
        _lib.check(L.fe_incidence_count(elem_type, nelem, nnode, _ptr(conn), _ptr(eptr), _ptr(ws), _stream(conn)))
        einc = torch.empty(nelem * nnpe, dtype=torch.int32, device=dev)
        cursor = torch.empty(max(1, nnode), dtype=torch.int32, device=dev)
        _lib.check(L.fe_incidence_fill(elem_type, nelem, nnode, _ptr(conn), _ptr(eptr), _ptr(cursor), _ptr(einc),
                                       _stream(conn)))
    return eptr, einc


def _adjacency_impl(conn: Tensor, eptr: Tensor, einc: Tensor, elem_type: int, between=None, after_fill=None):
    _chk_dev(conn, eptr, einc)
    nelem, nnpe = conn.shape
    nnode = eptr.numel() - 1
    dev = conn.device
    with torch.cuda.device(dev):
        L = _lib.lib()
        nptr = torch.empty(nnode + 1, dtype=torch.int32, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        ws = torch.empty(max(1, L.fe_scan_workspace_bytes(nnode + 1)), dtype=torch.uint8, device=dev)
        _lib.check(L.fe_adjacency_count(elem_type, nnode, _ptr(conn), _ptr(eptr), _ptr(einc), _ptr(nptr),
                                        _ptr(status), _ptr(ws), _stream(conn)))
        if between is not None:
            between()          # host work that may overlap the kernels queued so far (it must not touch this stream)
        tail = read_i32(nptr[nnode:nnode + 1], status)      # the one host sync of the symbolic phase
        if int(tail[1]) != 0:
            raise _lib.FeError(-4, "a node has more than %d distinct neighbours" % _lib.limits()["max_adj"])
        nnz_n = int(tail[0])
        nadj = torch.empty(max(1, nnz_n), dtype=torch.int32, device=dev)[:nnz_n]
        epos = torch.empty(max(1, nelem * nnpe * nnpe), dtype=torch.uint8, device=dev)[: nelem * nnpe * nnpe]
        _lib.check(L.fe_adjacency_fill(elem_type, nnode, _ptr(conn), _ptr(eptr), _ptr(einc), _ptr(nptr), _ptr(nadj),
                                       _ptr(epos), _stream(conn)))
        if after_fill is not None:
            after_fill(eptr, einc, nptr)       # eptr / einc / nptr are complete (the sync above), nadj / epos are being written
    return nptr, nadj, epos


@torch.library.custom_op("feprogram::adjacency", mutates_args=())
def adjacency(conn: Tensor, eptr: Tensor, einc: Tensor, elem_type: int) -> Tuple[Tensor, Tensor, Tensor]:
    """(nptr[nnode+1] int32, nadj[nnzN] int32, epos[nelem*nnpe*nnpe] uint8).  Syncs once to size nadj."""
    return _adjacency_impl(conn, eptr, einc, elem_type)


@torch.library.custom_op("feprogram::csr_expand", mutates_args=())
def csr_expand(nptr: Tensor, nadj: Tensor, ndof: int, idx64: bool) -> Tuple[Tensor, Tensor]:
    """(row_ptr[ndof*nnode+1] int32|int64, col_idx[nnz] int32): the scipy-compatible CSR pattern."""
    _chk_dev(nptr, nadj)
    nnode = nptr.numel() - 1
    dev = nptr.device
    with torch.cuda.device(dev):
        nnz = ndof * ndof * nadj.numel()
        row_ptr = torch.empty(ndof * nnode + 1, dtype=torch.int64 if idx64 else torch.int32, device=dev)
        col_idx = torch.empty(max(1, nnz), dtype=torch.int32, device=dev)[:nnz]
        _lib.check(_lib.lib().fe_csr_expand(ndof, nnode, _ptr(nptr), _ptr(nadj), int(idx64), _ptr(row_ptr),
                                            _ptr(col_idx), _stream(nptr)))
    return row_ptr, col_idx


@torch.library.custom_op("feprogram::assemble", mutates_args=("vals",))
def assemble(conn: Tensor, coords: Tensor, eptr: Tensor, einc: Tensor, epos: Tensor, nptr: Tensor, max_deg: int,
             vals: Tensor, elem_type: int, op: int) -> None:
    """Numeric assembly into vals (every entry written once, deterministic).  fe_assemble."""
    _chk_dev(conn, coords, eptr, einc, epos, nptr, vals)
    _want(conn, torch.int32, "conn")
    _want(coords, torch.float64, "coords")
    _want(vals, torch.float64, "vals")
    nelem = conn.shape[0]
    nnode = nptr.numel() - 1
    with torch.cuda.device(conn.device):
        _lib.check(_lib.lib().fe_assemble(elem_type, op, nelem, nnode, _ptr(conn), _ptr(coords), _ptr(eptr),
                                          _ptr(einc), _ptr(epos), _ptr(nptr), max_deg, _ptr(vals), _stream(conn)))


@torch.library.custom_op("feprogram::bc_rhs", mutates_args=("rhs", "dirmask"))
def bc_rhs(dir_dofs: Tensor, neu_dofs: Tensor, neu_value: float, rhs: Tensor, dirmask: Tensor) -> None:
    """rhs <- 0; rhs[dir] <- 0; rhs[neu] <- value (Neumann last); dirmask <- 1 on Dirichlet dofs."""
    _chk_dev(dir_dofs, neu_dofs, rhs, dirmask)
    _want(dir_dofs, torch.int64, "dir_dofs")
    _want(neu_dofs, torch.int64, "neu_dofs")
    _want(dirmask, torch.uint8, "dirmask")
    with torch.cuda.device(rhs.device):
        _lib.check(_lib.lib().fe_bc_rhs(rhs.numel(), _ptr(dir_dofs), dir_dofs.numel(), _ptr(neu_dofs),
                                        neu_dofs.numel(), float(neu_value), _ptr(rhs), _ptr(dirmask), _stream(rhs)))


@torch.library.custom_op("feprogram::edge_load", mutates_args=("rhs",))
def edge_load(edges: Tensor, coords: Tensor, tx: float, ty: float, dirmask: Tensor, rhs: Tensor) -> None:
    """rhs += consistent load of the traction (tx, ty) on the boundary edges (skipping Dirichlet dofs)."""
    _chk_dev(edges, coords, dirmask, rhs)
    _want(edges, torch.int32, "edges")
    _want(coords, torch.float64, "coords")
    _want(dirmask, torch.uint8, "dirmask")
    if edges.numel() == 0:
        return
    with torch.cuda.device(rhs.device):
        _lib.check(_lib.lib().fe_edge_load(edges.shape[1], edges.shape[0], _ptr(edges), _ptr(coords), float(tx),
                                           float(ty), _ptr(dirmask), _ptr(rhs), _stream(rhs)))


@torch.library.custom_op("feprogram::apply_bc", mutates_args=("vals", "rhs"))
def apply_bc(mode: int, row_ptr: Tensor, col_idx: Tensor, vals: Tensor, rhs: Tensor, dirmask: Tensor) -> None:
    """In-place Dirichlet treatment of the matrix: FE_BC_ROWS (reference) or FE_BC_SYMMETRIC (lift)."""
    _chk_dev(row_ptr, col_idx, vals, rhs, dirmask)
    with torch.cuda.device(vals.device):
        _lib.check(_lib.lib().fe_apply_bc(mode, row_ptr.numel() - 1, int(row_ptr.dtype == torch.int64), _ptr(row_ptr),
                                          _ptr(col_idx), _ptr(vals), _ptr(rhs), _ptr(dirmask), _stream(vals)))


@torch.library.custom_op("feprogram::spmv", mutates_args=("y",))
def spmv(row_ptr: Tensor, col_idx: Tensor, vals: Tensor, x: Tensor, y: Tensor, rowmask: Optional[Tensor]) -> None:
    """y <- A x (rows with rowmask != 0 give 0).  fe_spmv."""
    _chk_dev(row_ptr, col_idx, vals, x, y, rowmask)
    _want(vals, torch.float64, "vals")
    _want(x, torch.float64, "x")
    with torch.cuda.device(vals.device):
        _lib.check(_lib.lib().fe_spmv(row_ptr.numel() - 1, col_idx.numel(), int(row_ptr.dtype == torch.int64), _ptr(row_ptr),
                                      _ptr(col_idx), _ptr(vals), _ptr(x), _ptr(y), _ptr(rowmask), _stream(vals)))


@torch.library.custom_op("feprogram::spmv_nodal", mutates_args=("y",))
def spmv_nodal(nptr: Tensor, nadj: Tensor, vals: Tensor, x: Tensor, y: Tensor, rowmask: Optional[Tensor]) -> None:
    """y <- A x through the node-level pattern (2-dof matrices in this library's value layout).  fe_spmv_nodal."""
    _chk_dev(nptr, nadj, vals, x, y, rowmask)
    with torch.cuda.device(vals.device):
        _lib.check(_lib.lib().fe_spmv_nodal(nptr.numel() - 1, _ptr(nptr), _ptr(nadj), _ptr(vals), _ptr(x), _ptr(y),
                                            _ptr(rowmask), _stream(vals)))


@torch.library.custom_op("feprogram::csr_diagonal", mutates_args=())
def csr_diagonal(row_ptr: Tensor, col_idx: Tensor, vals: Tensor) -> Tensor:
    _chk_dev(row_ptr, col_idx, vals)
    n = row_ptr.numel() - 1
    with torch.cuda.device(vals.device):
        d = torch.empty(n, dtype=torch.float64, device=vals.device)
        _lib.check(_lib.lib().fe_csr_diagonal(n, int(row_ptr.dtype == torch.int64), _ptr(row_ptr), _ptr(col_idx),
                                              _ptr(vals), _ptr(d), _stream(vals)))
    return d


@torch.library.custom_op("feprogram::pcg", mutates_args=("x",))
def pcg(row_ptr: Tensor, col_idx: Tensor, vals: Tensor, rhs: Tensor, dirmask: Optional[Tensor], x: Tensor,
        rtol: float, atol: float, maxit: int, check_every: int, nptr: Optional[Tensor] = None,
        nadj: Optional[Tensor] = None) -> Tensor:
    """Jacobi-PCG with Dirichlet dofs held at rhs[D]; returns CPU tensor [iters, relres, ||r0||, stop] with
    stop = 1 converged, 0 iteration limit, -2 CG breakdown (matrix not SPD on the free dofs / NaN).
    nptr/nadj: node-level pattern of a 2-dof matrix (node-block SpMV inside the iteration)."""
    _chk_dev(row_ptr, col_idx, vals, rhs, dirmask, x, nptr, nadj)
    n = row_ptr.numel() - 1
    info = np.zeros(4, dtype=np.float64)
    with torch.cuda.device(vals.device):
        L = _lib.lib()
        work = torch.empty(L.fe_pcg_workspace_bytes(n), dtype=torch.uint8, device=vals.device)
        _lib.check(L.fe_pcg(n, col_idx.numel(), int(row_ptr.dtype == torch.int64), _ptr(row_ptr), _ptr(col_idx), _ptr(vals), _ptr(rhs),
                            _ptr(dirmask), _ptr(x), float(rtol), float(atol), int(maxit), int(check_every), _ptr(work),
                            info.ctypes.data_as(C.c_void_p), _ptr(nptr), _ptr(nadj), _stream(vals)))
    return torch.from_numpy(info)


# --------------------------------------------------------------------------------------------
# symbolic phase driver
# --------------------------------------------------------------------------------------------
@dataclass
class Pattern:
    """Everything the one-time symbolic phase produces (device tensors, reference numbering)."""
    elem_type: int
    ndof: int
    nnode: int
    nelem: int
    eptr: Tensor
    einc: Tensor
    nptr: Tensor
    nadj: Tensor
    epos: Tensor
    _row_ptr: Optional[Tensor] = None   # scipy-compatible dof-level CSR pattern: expanded eagerly by symbolic(), or on the
    _col_idx: Optional[Tensor] = None   # first read of row_ptr / col_idx (symbolic(..., csr="lazy"))
    _max_deg: Optional[int] = None
    lattice: Optional[tuple] = None    # (Morton keys from a lattice numbering, status, event) -- fe_lattice_keys
    csr_event: Optional[object] = None                 # set when row_ptr / col_idx were expanded on a side stream

    def csr_ready(self) -> None:
        """Orders the current stream after the expansion of row_ptr / col_idx (no-op unless they came from a side stream)."""
        ev, self.csr_event = self.csr_event, None
        if ev is not None:
            torch.cuda.current_stream(self.nptr.device).wait_event(ev)

    def _expand(self) -> None:
        if self._col_idx is None:       # lazy pattern: expand now, on the current stream (nptr / nadj are complete on it)
            self._row_ptr, self._col_idx = torch.ops.feprogram.csr_expand(self.nptr, self.nadj, self.ndof, self.nnz >= 2 ** 31)
        self.csr_ready()

    @property
    def row_ptr(self) -> Tensor:
        """Row pointers of the dof-level CSR matrix (int32, int64 from 2^31 non-zeros): what scipy's K.tocsr() holds."""
        self._expand()
        return self._row_ptr

    @property
    def col_idx(self) -> Tensor:
        """Column indices of the dof-level CSR matrix (int32), rows in ascending column order."""
        self._expand()
        return self._col_idx

    @property
    def max_deg(self) -> int:
        """Largest number of distinct neighbours of a node (sizes the row kernel's tile); read back on first use."""
        if self._max_deg is None:
            self._max_deg = int((self.nptr[1:] - self.nptr[:-1]).max().item()) if self.nnode > 0 else 0
        return self._max_deg

    @property
    def nnz(self) -> int:
        return self.ndof * self.ndof * self.nadj.numel()

    @property
    def nrows(self) -> int:
        return self.ndof * self.nnode


def symbolic(conn: Tensor, nnode: int, elem_type: int, ndof: int = 2, csr_stream=None, between=None,
             after_fill=None, csr: str = "eager") -> Pattern:
    """CSR pattern + element->slot map for a mesh (conn int32 0-based on the device).  ``csr_stream``: expand the
    scipy-compatible row_ptr / col_idx on that stream (concurrently with whatever follows on the current one); the
    caller orders later users with ``Pattern.csr_ready()``.  ``csr="lazy"``: do not expand them at all until something
    reads ``Pattern.row_ptr`` / ``col_idx`` -- the assembly kernels, the node-block SpMV and the PCG solve work on the
    node-level pattern (nptr, nadj), of which the dof-level one is a pure function (fe_csr_expand).  ``between(lattice)``: called once the first half of the
    symbolic kernels is queued and before the phase's host synchronisation -- the facade forms the tiles of a lattice
    numbering on another stream there (``tile_formation``), concurrently with those kernels.
    ``after_fill(eptr, einc, nptr)``: called right after the last symbolic kernel (fe_adjacency_fill) is queued; its
    arguments are complete by then, so the facade builds the first half of the tile plan on another stream while that
    kernel runs (``plan_layout_early``)."""
    nelem = conn.shape[0]
    lattice = None
    if elem_type == FE_QUAD4 and nelem > 0:
        # is the numbering a row-major lattice?  Then the tile plan needs no coordinates (fe_lattice_keys); the answer
        # is read back by tile_plan / tile_formation -- no extra synchronisation here
        with torch.cuda.device(conn.device):
            keys = torch.empty(nelem, dtype=torch.int32, device=conn.device)
            lstat = torch.empty(2, dtype=torch.int32, device=conn.device)
            _lib.check(_lib.lib().fe_lattice_keys(elem_type, nelem, _ptr(conn), _ptr(keys), _ptr(lstat), _stream(conn)))
            lattice = (keys, lstat, torch.cuda.current_stream(conn.device).record_event())
    eptr, einc = torch.ops.feprogram.incidence(conn, nnode, elem_type)
    hook = (lambda: between(lattice)) if between is not None else None
    nptr, nadj, epos = _adjacency_impl(conn, eptr, einc, elem_type, hook, after_fill)
    nnz = ndof * ndof * nadj.numel()
    ev = None
    if csr not in ("eager", "lazy"):
        raise ValueError("csr must be 'eager' or 'lazy'")
    if csr == "lazy":
        row_ptr = col_idx = None
    elif csr_stream is None:
        row_ptr, col_idx = torch.ops.feprogram.csr_expand(nptr, nadj, ndof, nnz >= 2 ** 31)
    else:
        main = torch.cuda.current_stream(conn.device)
        csr_stream.wait_event(main.record_event())
        with torch.cuda.stream(csr_stream):
            row_ptr, col_idx = torch.ops.feprogram.csr_expand(nptr, nadj, ndof, nnz >= 2 ** 31)
            ev = csr_stream.record_event()
        for t in (row_ptr, col_idx):
            t.record_stream(main)
        for t in (nptr, nadj):
            t.record_stream(csr_stream)
    return Pattern(elem_type, ndof, nnode, nelem, eptr, einc, nptr, nadj, epos, row_ptr, col_idx, lattice=lattice,
                   csr_event=ev)


# --------------------------------------------------------------------------------------------
# locality renumbering (solver-internal; the reference numbering stays the public one)
# --------------------------------------------------------------------------------------------
def numbering_span(conn: Tensor) -> float:
    """Median over the elements of (largest - smallest node id): ~ the row length of the lattice for a
    row-major numbering, ~ nnode / 2 for a random one."""
    if conn.shape[0] == 0:
        return 0.0
    c = conn[:: max(1, conn.shape[0] // 65536)].to(torch.int64)
    return float((c.max(dim=1).values - c.min(dim=1).values).to(torch.float64).median().item())


def locality_order(coords: Tensor) -> Tuple[Tensor, Tensor]:
    """(perm, order): Morton (Z-curve) order of the nodes; node `old` becomes node perm[old], order = perm^-1.
    Used only inside the solver: assembly tiles then own (nearly) contiguous CSR row ranges and the SpMV
    gathers of x stay in L1/L2, whatever the input numbering (BASELINE cfg5: shuffled node numbering)."""
    xy = coords[:, :2]
    lo = xy.min(dim=0).values
    span = (xy.max(dim=0).values - lo).clamp_min(1e-300)
    q = ((xy - lo) / span * 65535.0).to(torch.int64).clamp_(0, 65535)

    def spread(v):
        v = (v | (v << 8)) & 0x00FF00FF
        v = (v | (v << 4)) & 0x0F0F0F0F
        v = (v | (v << 2)) & 0x33333333
        v = (v | (v << 1)) & 0x55555555
        return v

    key = spread(q[:, 0]) | (spread(q[:, 1]) << 1)
    order = torch.argsort(key, stable=True)
    perm = torch.empty_like(order)
    perm[order] = torch.arange(order.numel(), device=order.device)
    return perm, order


# --------------------------------------------------------------------------------------------
# tile plan for the fast assembly kernel
# --------------------------------------------------------------------------------------------


@dataclass
class TilePlan:
    """One-time plan of fe_assemble_tiled (device tensors, tile-major; formats in csrc/fe_tiles.cu)."""
    tconn: Tensor           # int32 [ntiles, tile_rows, nnpe]: connectivity in tile order (own + halo rows)
    tdesc: Tensor           # int32 [ntiles, 8] (barrier-phased kernel) or [ntiles, 4] (warp-specialised kernel)
    tnode: Optional[Tensor]   # int32 [nslots, 2]                                              } barrier-phased
    tloc: Optional[Tensor]    # uint16 [items]                                                 } kernel only
    twork: Optional[Tensor]   # int32 (uint32 bit pattern) [work items, padded to 32 per tile] } (fe_assemble_tiled)
    tchunk: Optional[Tensor]  # int32 (uint32 bit pattern) [work items / 32]                   }
    ntiles: int
    stats: dict
    blobs: Optional[Tensor] = None    # uint8: per-tile record blobs  } warp-specialised kernel (fe_assemble_ws)
    truns: Optional[Tensor] = None    # int32 [runs, 4]               }

    @property
    def warp_specialised(self) -> bool:
        return self.blobs is not None


def _scan_into(src: Tensor, dst: Tensor, ws: Tensor):
    _lib.check(_lib.lib().fe_exclusive_scan_i32(_ptr(src), _ptr(dst), src.numel(), _ptr(ws), _stream(src)))


def _form_blocks(elem_type: int, nelem: int, te: int, keys: Tensor, imax: int):
    """Device-side tile formation on the current stream: counting sort of the elements by Morton block + per-block
    sort (csrc/fe_tile_form.cu).  One host synchronisation (the tile count).  Returns (ntiles, eperm, eloc,
    tile_estart) or None when a block overflows."""
    L = _lib.lib()
    dev = keys.device
    st = _stream(keys)
    shift = te.bit_length() - 1
    nb = max(1, (1 << (2 * max(1, imax.bit_length()))) >> shift)
    sb = nb + 1
    wsf = torch.empty(max(1, L.fe_scan_workspace_bytes(sb)), dtype=torch.uint8, device=dev)
    tabs = torch.empty((4, sb), dtype=torch.int32, device=dev)          # hist, tpb, ebase, tbase
    cursor_b = torch.empty(nb, dtype=torch.int32, device=dev)
    tmp = torch.empty(nelem, dtype=torch.int32, device=dev)
    fstat = torch.empty(1, dtype=torch.int32, device=dev)
    _lib.check(L.fe_tile_form_count(elem_type, nelem, nb, _ptr(keys), _ptr(tabs[0]), _ptr(tabs[1]),
                                    _ptr(tabs[2]), _ptr(tabs[3]), _ptr(cursor_b), _ptr(tmp), _ptr(fstat),
                                    _ptr(wsf), st))
    ft = read_i32(tabs[3, nb:nb + 1], fstat)                       # host sync: tile count
    if int(ft[1]) != 0:
        return None
    ntiles = int(ft[0])
    eperm = torch.empty(nelem, dtype=torch.int32, device=dev)
    eloc = torch.empty(nelem, dtype=torch.int32, device=dev)
    tile_estart = torch.empty(ntiles + 1, dtype=torch.int32, device=dev)
    _lib.check(L.fe_tile_form_sort(elem_type, nb, _ptr(tabs[2]), _ptr(tabs[3]), _ptr(tmp), _ptr(keys),
                                   _ptr(eperm), _ptr(eloc), _ptr(tile_estart), st))
    return ntiles, eperm, eloc, tile_estart


def _lattice_imax(w_lat: int, nnode: int) -> int:
    return max(w_lat - 1, (nnode + w_lat - 1) // w_lat)


def tile_formation(nelem: int, nnode: int, elem_type: int, lattice, stream) -> Optional[dict]:
    """Early tile formation for a lattice numbering, on ``stream``: element order, tile boundaries and the
    element -> (tile, row) map depend on the connectivity alone, so they can be built while the symbolic kernels run
    on the caller's stream (``symbolic(..., between=...)``).  Synchronises ``stream`` only.  Returns None when the
    numbering is not a lattice (tile_plan then takes its usual route)."""
    if lattice is None or nelem == 0:
        return None
    te = _lib.tile_params(elem_type)["tile_elems"]
    if te == 0:
        return None
    keys, lstat, ev = lattice
    main = torch.cuda.current_stream(keys.device)
    with torch.cuda.device(keys.device), torch.cuda.stream(stream):
        stream.wait_event(ev)
        keys.record_stream(stream)
        flag, w = read_i32(lstat)
        if flag != 0 or w < 2:
            return None
        res = _form_blocks(elem_type, nelem, te, keys, min(32767, _lattice_imax(w, nnode)))
        if res is None:
            return None
        done = stream.record_event()
    for t in res[1:]:
        t.record_stream(main)
    return dict(w=w, ntiles=res[0], eperm=res[1], eloc=res[2], tile_estart=res[3], event=done)


def tile_plan(conn: Tensor, coords: Tensor, pat: Pattern, morton: bool = True, kernel: str = "auto",
              coords_ready=None, formation: Optional[dict] = None, layout: Optional[dict] = None) -> Optional[TilePlan]:
    """Builds the tile plan; returns None when the mesh does not fit the tiled kernel's capacities
    (the caller then uses the row-tile kernel fe_assemble, which has none).  ``kernel``: "auto" = the
    warp-specialised kernel where it exists (Quad4 / Tri3) and its capacities hold, else the barrier-phased one;
    "ws" / "phased" force one of them (tests compare the two bit for bit).  ``coords_ready``: called before the
    coordinates are first read (the facade passes the wait for their upload); a mesh whose numbering is a row-major
    lattice (Pattern.lattice) gets its tiles from the numbering and never calls it.  ``formation`` / ``layout``: the
    results of ``tile_formation`` / ``plan_layout_early`` when the caller ran them ahead of time on a side stream.

    Elements are ordered along the Morton curve of a grid of one-element cells.  Strategy 1 ("blocks"): a
    tile = one aligned Z-block of 128 cells (exact 8x16 patches on structured meshes).  If that needs many
    more tiles than nelem/TE (irregular meshes: blocks with > TE elements get split, others are underfull)
    or overflows a capacity, strategy 2 ("pack") fills tiles with consecutive 16-cell sub-blocks up to TE
    elements; coarser cells are the last resort."""
    if kernel not in ("auto", "ws", "phased"):
        raise ValueError("kernel must be 'auto', 'ws' or 'phased'")
    if not morton:
        return _tile_plan(conn, coords, pat, False, 1.0, "blocks", kernel)
    te = _lib.tile_params(pat.elem_type)["tile_elems"]
    if te == 0 or pat.nelem == 0:
        return None
    ideal = (pat.nelem + te - 1) // te
    if layout is not None:                 # first half of the plan built ahead of time, on another stream (plan_layout_early)
        torch.cuda.current_stream(conn.device).wait_event(layout["event"])
        plan = _plan_records(conn, pat, layout, kernel)
        if plan is not None:
            return plan
    if formation is not None and pat.lattice is not None:       # tiles formed ahead of time (tile_formation)
        torch.cuda.current_stream(conn.device).wait_event(formation["event"])
        plan = _tile_plan(conn, coords, pat, True, 1.0, "blocks", kernel, lattice=(pat.lattice[0], formation["w"]),
                          formed=formation)
        if plan is not None:
            return plan
    elif pat.lattice is not None:
        flag, w = read_i32(pat.lattice[1])
        if flag == 0 and w >= 2:
            plan = _tile_plan(conn, coords, pat, True, 1.0, "blocks", kernel, lattice=(pat.lattice[0], w))
            if plan is not None:
                return plan
    if coords_ready is not None:
        coords_ready()
    best = _tile_plan(conn, coords, pat, True, 1.0, "blocks", kernel)
    if best is not None and best.ntiles <= 1.12 * ideal + 8:
        return best
    # irregular meshes: slightly smaller cells keep most Z-blocks at or below TE elements (a block with more is
    # split into a full and a nearly empty tile); packing sub-blocks and coarser cells are the fall-backs
    for mode, scale, good in (("blocks", 0.97, 1.08), ("blocks", 0.94, 1.10), ("pack", 1.0, 1.12), ("blocks", 0.9, 1.25),
                              ("pack", 1.25, 1.25), ("pack", 1.6, 1.25)):
        plan = _tile_plan(conn, coords, pat, True, scale, mode, kernel)
        if plan is not None and (best is None or plan.ntiles < best.ntiles):
            best = plan
        if best is not None and best.ntiles <= good * ideal + 8:
            break
    return best


def _tile_plan(conn: Tensor, coords: Tensor, pat: Pattern, morton: bool, cell_scale: float,
               mode: str, kernel: str = "auto", lattice=None, formed: Optional[dict] = None) -> Optional[TilePlan]:
    L = _lib.lib()
    params = _lib.tile_params(pat.elem_type)
    if params["halo_cap"] == 0 or pat.nelem == 0 or pat.nnode == 0:
        return None
    dev = conn.device
    te = params["tile_elems"]
    nelem, nnode = pat.nelem, pat.nnode
    mode_name = mode
    with torch.cuda.device(dev):
        st = _stream(conn)
        # ---- element order and tile boundaries ----
        idx = None
        keys = None
        if formed is not None:                           # tile_formation() has done this part on another stream
            ntiles, eperm, eloc, tile_estart = formed["ntiles"], formed["eperm"], formed["eloc"], formed["tile_estart"]
            mode_name = "lattice"
            formed = True
        elif lattice is not None:
            keys, w_lat = lattice                        # keys from the numbering: one cell per element, no coordinates
            imax_lat = _lattice_imax(w_lat, nnode)
            mode_name = "lattice"
            formed = False
        elif morton:
            formed = False
            # cell: holds ~ one element (Tri3: the two triangles of a quad).  Its AREA comes from the sampled mean
            # element area (distortion-proof: perturbing nodes changes extents, not the mean area), its aspect
            # ratio from the sampled mean extents.  All scalars come back in ONE device -> host transfer.
            sample = conn[:: max(1, nelem // 65536)][:65536].long()
            xy = coords[sample]                                               # (m, nnpe, 2)
            ext_e = (xy.amax(dim=1) - xy.amin(dim=1)).mean(dim=0)
            nc = 3 if pat.elem_type == FE_TRI3 else 4                         # corner nodes come first
            cx, cy = xy[:, :nc, 0], xy[:, :nc, 1]
            area = 0.5 * (cx * torch.roll(cy, -1, dims=1) - torch.roll(cx, -1, dims=1) * cy).sum(dim=1).abs().mean()
            box = torch.empty(4 * 593, dtype=torch.float64, device=dev)           # [0:4] result, rest scratch
            _lib.check(L.fe_bbox(nnode, _ptr(coords), _ptr(box[4:]), _ptr(box), st))
            sc = read_f64(box[:4], ext_e, area.reshape(1))
            lo_h = sc[0:2]
            ext = [max(sc[2] - sc[0], 1e-300), max(sc[3] - sc[1], 1e-300)]
            ext_e = [max(sc[4], 1e-300), max(sc[5], 1e-300)]
            cell_area = sc[6] * (2.0 if pat.elem_type == FE_TRI3 else 1.0)
            aspect = ext_e[0] / ext_e[1]
            size = [float(np.sqrt(cell_area * aspect)), float(np.sqrt(cell_area / aspect))]
            if not (np.isfinite(size[0]) and np.isfinite(size[1]) and min(size) > 0.0):
                size = [e * (0.7071 if pat.elem_type == FE_TRI3 else 1.0) for e in ext_e]
            cell = [max(size[k] * cell_scale, ext[k] / 32767.0, 1e-300) for k in range(2)]   # 15 bits per axis
            keys = torch.empty(nelem, dtype=torch.int32, device=dev)
            _lib.check(L.fe_morton_keys(pat.elem_type, nelem, _ptr(conn), _ptr(coords), lo_h[0], lo_h[1], cell[0],
                                        cell[1], _ptr(keys), st))
            imax_lat = None
        else:
            formed = False
        if keys is not None and mode == "blocks":
            # device-side formation: counting sort by Morton block + per-block sort (csrc/fe_tile_form.cu)
            imax = (min(32767, imax_lat) if imax_lat is not None else
                    max(min(32767, int(ext[0] / cell[0]) + 1), min(32767, int(ext[1] / cell[1]) + 1)))
            res = _form_blocks(pat.elem_type, nelem, te, keys, imax)
            if res is not None:
                ntiles, eperm, eloc, tile_estart = res
                formed = True
        if formed:
            pass
        elif morton:
            idx = torch.arange(nelem, dtype=torch.int64, device=dev)
            skeys, order = torch.sort(keys, stable=True)
            eperm = order.to(torch.int32)
            if mode == "pack":
                # consecutive 16-cell sub-blocks are packed into tiles of at most TE elements
                sub = skeys >> max(0, te.bit_length() - 1 - 3)                     # sub-blocks of TE/8 cells
                sflag = torch.ones(nelem, dtype=torch.bool, device=dev)
                sflag[1:] = sub[1:] != sub[:-1]
                sstart = torch.nonzero(sflag).flatten()
                ssize = torch.diff(torch.cat([sstart, torch.tensor([nelem], device=dev)]))
                m = int(ssize.max().item())
                if m > te // 2:
                    return None
                block = sstart[torch.cumsum(sflag, dim=0) - 1] // (te - m + 1)
                del sub, sflag, sstart, ssize
            else:
                block = skeys >> (te.bit_length() - 1)                        # aligned Z-blocks of TE cells
            del keys, skeys, order
        else:
            idx = torch.arange(nelem, dtype=torch.int64, device=dev)
            eperm = idx.to(torch.int32)
            block = idx // te
        if not formed:
            flag = torch.ones(nelem, dtype=torch.bool, device=dev)
            flag[1:] = block[1:] != block[:-1]
            seg_start = torch.nonzero(flag).flatten()[torch.cumsum(flag, dim=0) - 1]   # start of each element's block
            flag |= ((idx - seg_start) % te) == 0                                 # at most TE elements per tile
            tid = torch.cumsum(flag.to(torch.int32), dim=0, dtype=torch.int32) - 1
            ntiles = int(tid[-1].item()) + 1
            starts = torch.nonzero(flag).flatten()
            row = (idx - starts[tid.long()]).to(torch.int32)
            tile_estart = torch.cat([starts, torch.tensor([nelem], device=dev)]).to(torch.int32)
            # inside a tile order the elements by element id
            etile = torch.empty(nelem, dtype=torch.int32, device=dev)
            etile[eperm.long()] = tid
            eperm = torch.sort(etile, stable=True).indices.to(torch.int32)
            del etile
            eloc = torch.empty(nelem, dtype=torch.int32, device=dev)
            eloc[eperm.long()] = (tid << 8) | row          # tid / row are per sorted position
            del block, flag, seg_start, tid, starts, row, idx
        lay = _plan_layout(conn, pat.elem_type, nnode, pat.eptr, pat.einc, pat.nptr, ntiles, eperm, eloc, tile_estart)
    if lay is None:
        return None
    lay["info"].update(morton=morton, cell_scale=cell_scale, mode=mode_name)
    return _plan_records(conn, pat, lay, kernel)


def _plan_layout(conn: Tensor, elem_type: int, nnode: int, eptr: Tensor, einc: Tensor, nptr: Tensor, ntiles: int,
                 eperm: Tensor, eloc: Tensor, tile_estart: Tensor) -> Optional[dict]:
    """First half of the plan, on the current stream: node ownership, per-tile node / halo lists and offsets, the
    tile-ordered connectivity.  Needs the incidence and the node DEGREES (nptr) but neither nadj nor epos, so the facade
    runs it on a side stream while fe_adjacency_fill is still busy (``plan_layout_early``).  One host synchronisation
    (every size and capacity).  Returns None when a capacity of the tiled kernels is exceeded."""
    L = _lib.lib()
    params = _lib.tile_params(elem_type)
    hcap = params["halo_cap"]
    dev = conn.device
    if ntiles >= (1 << 23):
        return None
    with torch.cuda.device(dev):
        st = _stream(conn)
        s = ntiles + 1
        einv = eloc
        ws = torch.empty(max(1, L.fe_scan_workspace_bytes(s)), dtype=torch.uint8, device=dev)
        owner = torch.empty(nnode, dtype=torch.int32, device=dev)
        tile_ncount = torch.empty(s, dtype=torch.int32, device=dev)
        _lib.check(L.fe_tile_owner(nnode, ntiles, _ptr(eptr), _ptr(einc), _ptr(einv), _ptr(owner),
                                   _ptr(tile_ncount), st))
        tile_ptr = torch.empty((6, s), dtype=torch.int32, device=dev)     # nodes, items, halo, work, blob / 16, runs
        _scan_into(tile_ncount, tile_ptr[0], ws)
        cursor = torch.empty(ntiles, dtype=torch.int32, device=dev)
        tile_nodes = torch.empty(nnode, dtype=torch.int32, device=dev)
        noff = torch.empty(2 * nnode, dtype=torch.int32, device=dev)
        thalo = torch.empty((ntiles, hcap), dtype=torch.int32, device=dev)
        counts = torch.empty((5, s), dtype=torch.int32, device=dev)
        stats = torch.empty(8, dtype=torch.int32, device=dev)
        _lib.check(L.fe_tile_layout(elem_type, nnode, ntiles, _ptr(owner), _ptr(tile_ptr[0]), _ptr(eptr), _ptr(einc),
                                    _ptr(einv), _ptr(nptr), _ptr(cursor), _ptr(tile_nodes), _ptr(noff),
                                    _ptr(thalo), _ptr(counts), _ptr(stats), st))
        for r in range(5):
            _scan_into(counts[r], tile_ptr[r + 1], ws)
        tail = read_i32(tile_ptr[:, ntiles], stats)                         # one host sync: every size and capacity
        nslots, n_items, n_halo, n_work, blob16, nruns = (int(v) for v in tail[:6])
        max_nodes, max_items, max_halo, max_cnt = (int(v) for v in tail[6:10])
        max_work, max_blob, max_stage = int(tail[11]), int(tail[12]), int(tail[13])
        info = dict(max_nodes=max_nodes, max_items=max_items, max_halo=max_halo, max_cnt=max_cnt, max_work=max_work,
                    items=n_items, halo_elems=n_halo, work_items=n_work, ntiles=ntiles)
        if (max_nodes > params["max_nodes"] or max_items >= params["max_items"] - 1 or max_halo > hcap
                or max_work > params["max_work"]):
            return None
        tconn = torch.empty((ntiles, params["tile_rows"], conn.shape[1]), dtype=torch.int32, device=dev)
        _lib.check(L.fe_tile_conn(elem_type, ntiles, _ptr(conn), _ptr(eperm), _ptr(tile_estart), _ptr(thalo),
                                  _ptr(tile_ptr), _ptr(tconn), st))
    return dict(ntiles=ntiles, eloc=eloc, owner=owner, tile_ptr=tile_ptr, tile_nodes=tile_nodes, noff=noff, thalo=thalo,
                stats=stats, tconn=tconn, info=info, nslots=nslots, n_items=n_items, n_work=n_work, blob16=blob16,
                nruns=nruns, max_nodes=max_nodes, max_blob=max_blob, max_stage=max_stage)


def plan_layout_early(conn: Tensor, elem_type: int, nnode: int, eptr: Tensor, einc: Tensor, nptr: Tensor,
                      formation: Optional[dict], stream) -> Optional[dict]:
    """``_plan_layout`` of a lattice numbering on ``stream`` (the one ``tile_formation`` ran on), concurrently with
    whatever the caller's stream is doing -- the facade calls it right after queueing fe_adjacency_fill, whose inputs
    (incidence, node degrees) are complete by then (``symbolic(after_fill=...)`` runs behind the phase's host
    synchronisation).  Synchronises ``stream`` only.  Returns None when there is nothing formed or a capacity fails."""
    if formation is None:
        return None
    dev = conn.device
    main = torch.cuda.current_stream(dev)
    with torch.cuda.device(dev), torch.cuda.stream(stream):
        for t in (conn, eptr, einc, nptr):
            t.record_stream(stream)
        lay = _plan_layout(conn, elem_type, nnode, eptr, einc, nptr, formation["ntiles"], formation["eperm"],
                           formation["eloc"], formation["tile_estart"])
        if lay is None:
            return None
        lay["event"] = stream.record_event()
    for v in lay.values():
        if isinstance(v, Tensor):
            v.record_stream(main)
    lay["info"].update(morton=True, cell_scale=1.0, mode="lattice")
    return lay


def _plan_records(conn: Tensor, pat: Pattern, lay: dict, kernel: str) -> Optional[TilePlan]:
    """Second half of the plan: the per-tile records of the numeric kernel (needs nadj / epos)."""
    L = _lib.lib()
    dev = conn.device
    ntiles, einv, owner, tile_ptr, tile_nodes = lay["ntiles"], lay["eloc"], lay["owner"], lay["tile_ptr"], lay["tile_nodes"]
    noff, thalo, stats, tconn, info = lay["noff"], lay["thalo"], lay["stats"], lay["tconn"], lay["info"]
    nslots, n_items, n_work, blob16, nruns = lay["nslots"], lay["n_items"], lay["n_work"], lay["blob16"], lay["nruns"]
    max_nodes, max_blob, max_stage = lay["max_nodes"], lay["max_blob"], lay["max_stage"]
    with torch.cuda.device(dev):
        st = _stream(conn)
        wsp = _lib.tile_ws_params(pat.elem_type)
        # "auto": the warp-specialised kernel where it is the faster one (measured: Quad4 1.39 ms against 1.90 ms on cfg2;
        # Tri3's tiles are too small for its per-tile hand-offs: 1.51 ms against 1.25 ms), else the barrier-phased kernel
        # (one-dof patterns use the same plan: the kernel's SCALAR form stages 8-byte values and writes the runs with
        # plain coalesced stores)
        want_ws = kernel == "ws" or (kernel == "auto" and pat.elem_type == FE_QUAD4)
        if want_ws and wsp["blob_cap"] > 0:
            # ---- warp-specialised kernel: per-tile record blobs + runs of contiguous CSR rows ----
            wstats = torch.zeros(8, dtype=torch.int32, device=dev)
            info.update(max_blob=max_blob, max_stage=max_stage, runs=nruns, blob_bytes=16 * blob16)
            fits = (max_blob <= wsp["blob_cap"] and max_stage <= wsp["stage_cap"] and max_nodes <= wsp["max_nodes"]
                    and blob16 < 2 ** 31)
            if fits:
                blobs = torch.empty(max(16, 16 * blob16), dtype=torch.uint8, device=dev)
                truns = torch.empty((max(1, nruns), 4), dtype=torch.int32, device=dev)
                tdesc = torch.empty((ntiles, 4), dtype=torch.int32, device=dev)
                _lib.check(L.fe_tile_ws_build(pat.elem_type, nslots, ntiles, _ptr(tile_nodes), _ptr(owner), _ptr(tile_ptr),
                                              _ptr(noff), _ptr(thalo), _ptr(pat.eptr), _ptr(pat.einc), _ptr(pat.epos),
                                              _ptr(einv), _ptr(pat.nptr), _ptr(pat.nadj), _ptr(tile_ptr[4]), _ptr(tile_ptr[5]),
                                              _ptr(blobs), _ptr(truns), _ptr(tdesc), _ptr(wstats), st))
                if read_i32(wstats[4:5])[0] != 0:
                    return None
                info["kernel"] = "ws"
                return TilePlan(tconn, tdesc, None, None, None, None, ntiles, info, blobs, truns)
            if kernel == "ws":
                return None
        elif kernel == "ws":
            return None
        info["kernel"] = "phased"
        tnode = torch.empty((max(1, nslots), 2), dtype=torch.int32, device=dev)
        tloc = torch.zeros(max(2, n_items), dtype=torch.uint16, device=dev)
        twork = torch.zeros(max(32, n_work), dtype=torch.int32, device=dev)
        tchunk = torch.zeros(max(1, n_work // 32), dtype=torch.int32, device=dev)
        _lib.check(L.fe_tile_records(pat.elem_type, nslots, _ptr(tile_nodes), _ptr(owner), _ptr(tile_ptr), ntiles,
                                     _ptr(noff), _ptr(thalo), _ptr(pat.eptr), _ptr(pat.einc), _ptr(pat.epos),
                                     _ptr(einv), _ptr(pat.nptr), _ptr(pat.nadj), _ptr(tnode), _ptr(tloc),
                                     _ptr(twork), _ptr(tchunk), _ptr(stats), st))
        if read_i32(stats[4:5])[0] != 0:
            return None
        z = torch.zeros(ntiles, dtype=torch.int32, device=dev)
        tdesc = torch.stack([tile_ptr[0, :-1], tile_ptr[0, 1:] - tile_ptr[0, :-1],
                             tile_ptr[1, :-1], tile_ptr[1, 1:] - tile_ptr[1, :-1],
                             tile_ptr[3, :-1], tile_ptr[3, 1:] - tile_ptr[3, :-1],
                             tile_ptr[2, 1:] - tile_ptr[2, :-1], z], dim=1).contiguous()
    return TilePlan(tconn, tdesc, tnode, tloc, twork, tchunk, ntiles, info)


@torch.library.custom_op("feprogram::assemble_tiled", mutates_args=("vals",))
def assemble_tiled(tconn: Tensor, coords: Tensor, ntiles: int, tdesc: Tensor, tnode: Tensor, tloc: Tensor,
                   twork: Tensor, tchunk: Tensor, vals: Tensor, elem_type: int, op: int) -> None:
    """Numeric assembly, tiled kernel (same results as feprogram::assemble, bit for bit)."""
    _chk_dev(tconn, coords, tdesc, tnode, tloc, twork, tchunk, vals)
    _want(tconn, torch.int32, "tconn")
    _want(coords, torch.float64, "coords")
    _want(vals, torch.float64, "vals")
    with torch.cuda.device(vals.device):
        _lib.check(_lib.lib().fe_assemble_tiled(elem_type, op, ntiles, _ptr(tconn), _ptr(coords), _ptr(tdesc),
                                                _ptr(tnode), _ptr(tloc), _ptr(twork), _ptr(tchunk), _ptr(vals),
                                                _stream(vals)))


@torch.library.custom_op("feprogram::assemble_ws", mutates_args=("vals",))
def assemble_ws(tconn: Tensor, coords: Tensor, ntiles: int, tdesc: Tensor, blobs: Tensor, truns: Tensor, vals: Tensor,
                elem_type: int, op: int) -> None:
    """Numeric assembly, warp-specialised tiled kernel (same results as feprogram::assemble, bit for bit)."""
    _chk_dev(tconn, coords, tdesc, blobs, truns, vals)
    _want(tconn, torch.int32, "tconn")
    _want(coords, torch.float64, "coords")
    _want(vals, torch.float64, "vals")
    with torch.cuda.device(vals.device):
        _lib.check(_lib.lib().fe_assemble_ws(elem_type, op, ntiles, _ptr(tconn), _ptr(coords), _ptr(tdesc), _ptr(blobs),
                                             _ptr(truns), _ptr(vals), _stream(vals)))


def assemble_values(conn: Tensor, coords: Tensor, pat: Pattern, plan: Optional[TilePlan], vals: Tensor, op: int = 0):
    """Numeric assembly with whichever kernel the plan allows.  A one-dof pattern (``symbolic(..., ndof=1)``) takes
    FE_OP_POISSON and nothing else; two-dof patterns take the other operators."""
    if (pat.ndof == 1) != (op == FE_OP_POISSON):
        raise ValueError("operator %d does not match a pattern with %d dof per node" % (op, pat.ndof))
    if vals.numel() != pat.nnz:
        raise ValueError("vals must hold %d entries" % pat.nnz)
    if plan is not None and plan.warp_specialised:
        torch.ops.feprogram.assemble_ws(plan.tconn, coords, plan.ntiles, plan.tdesc, plan.blobs, plan.truns, vals,
                                        pat.elem_type, op)
    elif plan is not None:
        hint = 256 if plan.stats.get("mode") == "pack" else 0       # FE_HINT_IRREGULAR_TILES: schedule only
        torch.ops.feprogram.assemble_tiled(plan.tconn, coords, plan.ntiles, plan.tdesc, plan.tnode, plan.tloc,
                                           plan.twork, plan.tchunk, vals, pat.elem_type, op | hint)
    else:
        torch.ops.feprogram.assemble(conn, coords, pat.eptr, pat.einc, pat.epos, pat.nptr, pat.max_deg, vals,
                                     pat.elem_type, op)
