"""Drop-in for the reference's ``elements`` module on the hot path (B200 / CUDA only).

Same classes, method names, argument meaning and call order as /root/reference/elements.py
(``Elem`` :9-38, ``FemProblem`` :40-243) as driven by main.py:17-40::

    fem = FemProblem(conectivity, coordinates)      # 1-based tags, (nnode,2) fp64 coordinates
    fem.setMatrix()
    fem.setBoundaryConditions(bcNodeTags)
    fem.assemble()                                  # GPU: symbolic (once per mesh) + numeric assembly + BCs
    U = fem.solveProblem()                          # GPU: Jacobi-PCG; returns ndarray (2*nnode,)

What differs, deliberately:
  * the arithmetic runs in hand-written sm_100a kernels behind the C ABI (include/feprogram_b200.h)
    via torch custom ops (feprogram_b200/ops.py); without a CUDA device every hot call raises;
  * ``K`` / ``brhs`` live on the GPU; the ``scipy.sparse`` objects the reference exposes
    (``fem.K`` csc_matrix, ``fem.brhs`` csc (2N x 1)) are materialised lazily on attribute access;
  * ``solveProblem`` is Jacobi-PCG on the SPD system equivalent to the reference's row-only
    Dirichlet treatment instead of SuperLU (agreement <= 1e-8 relative L2; tests/test_gpu_parity.py).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from .tools.interpFunctions import funH4, funH9, funHrs4, funHrs9

__all__ = ["Elem", "FemProblem"]


class Elem:
    """Element descriptor (elements.py:9-38).  Host-side helpers kept for API parity; the per-element
    arithmetic the assembly uses is csrc/fe_element.cuh."""

    def __init__(self, nodes, elemType):
        self.elemType = elemType
        self.nnodesloc = nodes

    def _gradients(self, dHrs, J):
        return np.linalg.inv(J) * dHrs

    def funB(self, dHrs, J):
        """4 x 2n strain operator: rows [dx on u, dy on v, dy on u + dx on v, 0] (elements.py:15-23)."""
        n = self.nnodesloc
        Dh = np.asarray(self._gradients(np.asmatrix(dHrs), np.asmatrix(J)))
        B = np.zeros((4, 2 * n))
        B[0, 0::2] = Dh[0]
        B[1, 1::2] = Dh[1]
        B[2, 0::2] = Dh[1]
        B[2, 1::2] = Dh[0]
        return np.matrix(B)

    def getHgrad(self, dHrs, J):
        """4 x 2n velocity-gradient operator (elements.py:25-32); unused by the hot path."""
        n = self.nnodesloc
        Dh = np.asarray(self._gradients(np.asmatrix(dHrs), np.asmatrix(J)))
        G = np.zeros((4, 2 * n))
        G[0:2, 0::2] = Dh
        G[2:4, 1::2] = Dh
        return np.matrix(G)

    def getpos(self, nodesTagList, coordinates):
        """n x 2 matrix of the element's node coordinates, 0-based tags (elements.py:34-38)."""
        idx = np.asarray(nodesTagList, dtype=np.int64)
        return np.matrix(np.asarray(coordinates)[idx, :2])


_ELEM_CODES = {"Quad4": 4, "Quad9": 9}

_SIDE_STREAMS = {}


def side_streams(dev):
    """(upload stream, boundary-data stream, CSR-expansion stream) of a device, shared by every problem on it.
    PyTorch's caching allocator keeps one pool per stream: a problem that made its own streams would strand the
    blocks it allocated there (coordinates, rhs) when it dies, and every new problem would pay fresh cudaMallocs."""
    import torch
    key = (dev.type, dev.index)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = (torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev))
    return _SIDE_STREAMS[key]


class FemProblem:
    def __init__(self, conectivity, coordinates, *, device=None, operator="reference", renumber=None, trace=False):
        """conectivity: sequence of equal-length 1-D integer arrays of 1-BASED node tags (or a 2-D
        ndarray); coordinates: float64 (nnode, >=2), row i <-> tag i+1 (elements.py:41-58).

        Keyword-only extensions: ``device`` (torch device, default current CUDA device) and
        ``operator`` ("reference" = un-weighted B^T B detJ as elements.py:162, "weighted" = times gpWei,
        "mass" = sum_g gpWei H^T H detJ from the reference's H / gpWei tables, on the same CSR pattern,
        "poisson" = the scalar operator with ONE dof per node, sum_g grad N_a . grad N_b detJ = ``K[0::2, 0::2]`` of
        the reference matrix, bit for bit: dof = node, bottom side Dirichlet, left side nodal load);
        ``renumber``: solve on a Morton-renumbered copy of the mesh (None = only when the input numbering has no
        locality, e.g. BASELINE cfg5's shuffled tags).  Everything visible -- dof numbering, ``K``, ``brhs``,
        the solution -- stays in the reference's numbering; ``K`` is then assembled on demand in that numbering.
        ``trace``: record a CUDA event + host time stamp at every stage boundary of assemble() (no synchronisation is
        added); ``stage_times()`` returns them -- the device-side counterpart of the reference's timing log lines
        (main.py:14-46).
        """
        self.nnode = coordinates.shape[0]
        self.nelem = len(conectivity)
        self.elem = Elem(4, "Quad4") if len(conectivity[0]) == 4 else Elem(9, "Quad9")  # elements.py:55
        self.conectivity = conectivity
        self.coordinates = coordinates
        self._device = device
        self._trace = [] if trace else None
        self._renumber = renumber
        self._operator = {"reference": 0, "weighted": 1, "mass": 2, "poisson": 3}[operator]
        self.ndof = 1 if operator == "poisson" else 2         # dofs per node (the reference: 2, elements.py:148)
        self._dev = None        # device-side state, created on first assemble()
        self._K_host = None
        self._brhs_host = None
        self._assembled = False
        self.solver_info = None
        self.rtol = 1e-10
        self.maxit = None

    # ------------------------------------------------------------------ tracing
    def _mark(self, name, stream=None):
        if self._trace is None:
            return
        import time
        import torch
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(stream if stream is not None else torch.cuda.current_stream(self._torch_device()))
        self._trace.append((name, ev, time.perf_counter()))

    def stage_times(self):
        """{stage: (ms on the device, ms on the host)} since the first mark; synchronises the device."""
        import torch
        if not self._trace:
            return {}
        torch.cuda.synchronize(self._torch_device())
        _, ev0, t0 = self._trace[0]
        return {name: (round(ev0.elapsed_time(ev), 3), round(1e3 * (t - t0), 3)) for name, ev, t in self._trace}

    # ------------------------------------------------------------------ setMatrix (elements.py:60-92)
    def setMatrix(self):
        et = self.elem.elemType
        if et == "Quad9":
            gpsList = [-0.774, 0, 0.774]
            gpsWei = [0.55555, 0.88888, 0.55555]
            self.H = self.quadH(gpsList, funH9, gpsWei)
            self.Hrs = self.quadHrs(gpsList, funHrs9, gpsWei)
        elif et == "Quad4":
            gpsList = [-0.5773, 0.5773]
            gpsWei = [1, 1]
            self.H = self.quadH(gpsList, funH4, gpsWei)
            self.Hrs = self.quadHrs(gpsList, funHrs4, gpsWei)
        else:
            raise Exception("Invalid elemType defined you must use Quad4 or Quad9")
        # reset, like the reference's fresh empty lil K / brhs
        self._K_host = None
        self._brhs_host = None
        self._assembled = False
        if self._dev is not None:
            self._dev.pop("rhs", None)

    def quadH(self, gpoints, formFunction, gpweights=None):
        """Shape values on the tensor Gauss grid, r outer / s inner (elements.py:94-113)."""
        return [formFunction(r, s) for r in gpoints for s in gpoints]

    def quadHrs(self, gpoints, formFunction, gpweights=None):
        """Shape derivatives on the tensor Gauss grid + self.gpWei (elements.py:115-140)."""
        self.gpWei = [wr * ws for wr in gpweights for ws in gpweights]
        return [formFunction(r, s) for r in gpoints for s in gpoints]

    # ------------------------------------------------------- boundary conditions (elements.py:202-229)
    def setBoundaryConditions(self, bcNodesList, *, verbose=True, mode="reference", traction=(1.0, 1.0)):
        """bcNodesList = 4 iterables of 1-based tags [bottom, right, top, left]; bottom -> Dirichlet
        x,y; left -> nodal load ("Neumann") x,y.  Prints both dof sets like the reference unless
        ``verbose=False``.

        ``mode="reference"`` keeps the reference's semantics exactly (row-only Dirichlet, nodal load 1.0 per
        dof written last so it overrides the Dirichlet value at the shared corner; elements.py:185-197).
        ``mode="corrected"``: Dirichlet wins at shared nodes, the load is the consistent
        edge integral of the constant ``traction`` over the left side (mesh-size independent), and ``K`` is
        the symmetrically eliminated matrix."""
        if mode not in ("reference", "corrected"):
            raise ValueError("mode must be 'reference' or 'corrected'")
        if mode == "corrected" and self.ndof != 2:
            raise ValueError("mode='corrected' is defined for the two-dof operators")
        self._bc_mode = mode
        self._traction = (float(traction[0]), float(traction[1]))
        self._load_edges = None
        if mode == "corrected":
            from . import mesh
            self._load_edges = mesh.boundary_edges(self.conectivity, bcNodesList[3])
        dofSetDir = [[0, 1], [], [], []]
        dofSetNeu = [[], [], [], [0, 1]]
        dirDof, neuDof = set(), set()
        nd = self.ndof
        for enu, nodes in enumerate(bcNodesList):
            tags = np.fromiter((int(t) for t in nodes), dtype=np.int64)
            for k in dofSetDir[enu]:
                if k < nd:
                    dirDof.update(((tags - 1) * nd + k).tolist())
            for m in dofSetNeu[enu]:
                if m < nd:
                    neuDof.update(((tags - 1) * nd + m).tolist())
        if verbose:
            print(dirDof)
            print(neuDof)
        if mode == "corrected":
            neuDof -= dirDof
        self.dofDir = list(dirDof)
        self.dofNeu = list(neuDof)

    # --------------------------------------------------------------- small helpers (elements.py:143-163)
    def getRowData(self, elemNodesTag):
        dofs = [(int(x) - 1) * self.ndof + y for x in elemNodesTag for y in range(self.ndof)]
        m = len(dofs)
        return [d for d in dofs for _ in range(m)], dofs * m

    def getKe(self, elemNodeTags):
        """Element matrix of one element from 0-based tags, computed by the CUDA kernel (fe_element_matrices)."""
        import torch
        dev = self._torch_device()
        tags = torch.as_tensor(np.asarray(elemNodeTags, dtype=np.int64).astype(np.int32)[None, :], device=dev)
        st = self._state()
        ke = torch.ops.feprogram.element_matrices(tags.contiguous(), st["coords"], _ELEM_CODES[self.elem.elemType],
                                                  self._operator)
        return np.matrix(ke[0].cpu().numpy())

    def getLocalVec(self, ind, Vec):
        return Vec[ind].reshape((2, 4))

    def initialVector(self):
        return np.ones(self.nnode * self.ndof)

    def assembleMatrix(self, vec):
        return 0

    # ------------------------------------------------------- post-processing
    def _field(self, U):
        import torch
        st = self._state()
        u = torch.from_numpy(np.ascontiguousarray(np.asarray(U, dtype=np.float64).reshape(-1))).to(st["coords"].device)
        if self.ndof != 2 or u.numel() != 2 * self.nnode:
            raise ValueError("U must hold 2 values per node (post-processing is defined for the two-dof operators)")
        return st, u

    def gaussGradients(self, U):
        """(nelem, ngauss, 4): Elem.getHgrad(Hrs[g], J_g) * u_e = (du/dx, du/dy, dv/dx, dv/dy) at every Gauss
        point of every element (elements.py:25-32), computed by fe_gauss_gradients."""
        import torch
        st, u = self._field(U)
        return torch.ops.feprogram.gauss_gradients(st["conn"], st["coords"], u,
                                                   _ELEM_CODES[self.elem.elemType]).cpu().numpy()

    def nodalStrain(self, U):
        """(nnode, 3): nodal means of (du/dx, dv/dy, du/dy + dv/dx) -- the "Tensiones" array of the reference's
        writer (python_to_xml.py:19-22), computed by fe_nodal_strain."""
        import torch
        from . import ops
        st, u = self._field(U)
        et = _ELEM_CODES[self.elem.elemType]
        if "pattern" in st and st["perm"] is None:
            eptr, einc = st["pattern"].eptr, st["pattern"].einc
        else:
            eptr, einc = torch.ops.feprogram.incidence(st["conn"], self.nnode, et)
        return torch.ops.feprogram.nodal_strain(st["conn"], st["coords"], u, eptr, einc, et).cpu().numpy()

    # ------------------------------------------------------------------------------- device state
    def _torch_device(self):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("feprogram_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        return torch.device(self._device) if self._device is not None else torch.device("cuda", torch.cuda.current_device())

    def _state(self, defer_coords=False):
        """Uploads connectivity (int32, 0-based) and coordinates once.  The coordinates travel on a side stream
        behind the connectivity so that the first part of the symbolic phase (which only needs the
        connectivity) overlaps their transfer; ``defer_coords`` leaves the wait to the caller (assemble)."""
        if self._dev is not None:
            if not defer_coords:
                self._coords_ready()
            return self._dev
        import torch
        from . import ops  # noqa: F401  (registers torch.ops.feprogram.*)
        dev = self._torch_device()
        conn = self.conectivity
        if not isinstance(conn, np.ndarray) or conn.ndim != 2:
            conn = np.stack([np.asarray(c) for c in conn])
        if conn.dtype.kind not in "iu" or conn.dtype.itemsize not in (4, 8):
            conn = conn.astype(np.int64)
        if conn.dtype.kind == "u":                       # gmsh hands out uint64 tags; same bits as int64
            conn = conn.view(np.int64 if conn.dtype.itemsize == 8 else np.int32)
        xy = np.asarray(self.coordinates, dtype=np.float64)
        if xy.shape[1] != 2 or not xy.flags.c_contiguous:
            xy = np.ascontiguousarray(xy[:, :2])
        main = torch.cuda.current_stream(dev)
        self._mark("start", main)
        # boundary dof sets first (tiny): a small copy issued later would queue behind the 268 MB coordinate upload on the
        # host->device copy engine and hold back everything that depends on the right-hand side
        pre = {}
        if getattr(self, "dofDir", None) is not None and getattr(self, "dofNeu", None) is not None:
            pre["dof_sets"] = (self.dofDir, self.dofNeu,
                               torch.as_tensor(np.asarray(self.dofDir, dtype=np.int64), device=dev),
                               torch.as_tensor(np.asarray(self.dofNeu, dtype=np.int64), device=dev))
            pre["dof_event"] = main.record_event()
        tags = torch.from_numpy(np.ascontiguousarray(conn)).to(dev, non_blocking=True)
        self._mark("conn uploaded", main)
        side = side_streams(dev)[0]
        side.wait_event(main.record_event())          # one host->device transfer at a time: connectivity first
        with torch.cuda.stream(side):
            coords_d = torch.from_numpy(xy).to(dev, non_blocking=True)
            self._mark("coords uploaded", side)
        coords_d.record_stream(main)
        st = {
            "conn": torch.ops.feprogram.tags_to_conn(tags),                       # elements.py:172 `elem-1`
            "coords": coords_d,
            "coords_event": side.record_event(),
        }
        st.update(pre)
        self._dev = st
        ren = self._renumber
        if ren is None:      # a row-major lattice spans ~sqrt(nnode) ids per element, a shuffled one ~nnode/2
            ren = self.nnode > 100000 and ops.numbering_span(st["conn"]) > 64.0 * np.sqrt(self.nnode)
        st["conn_s"], st["coords_s"], st["perm"] = st["conn"], st["coords"], None
        if ren or not defer_coords:
            self._coords_ready()
        if ren:
            perm, order = ops.locality_order(st["coords"])
            st["perm"] = perm                                                     # reference node -> solver node
            st["conn_s"] = perm[st["conn"].to(torch.int64)].to(torch.int32).contiguous()
            st["coords_s"] = st["coords"][order].contiguous()
        self._dev = st
        return self._dev

    def _coords_ready(self):
        """Makes the current stream wait for the coordinate upload (once)."""
        import torch
        ev = self._dev.pop("coords_event", None)
        if ev is not None:
            torch.cuda.current_stream(self._dev["conn"].device).wait_event(ev)

    def _solver_dofs(self, dofs):
        """Reference dof ids -> dof ids of the (possibly renumbered) solver mesh, as a device tensor."""
        import torch
        st = self._dev
        pre = st.get("dof_sets")
        if pre is not None and dofs is pre[0]:            # uploaded ahead of the mesh by _state()
            d = pre[2]
        elif pre is not None and dofs is pre[1]:
            d = pre[3]
        else:
            d = torch.as_tensor(np.asarray(dofs, dtype=np.int64), device=st["conn"].device)
        if st["perm"] is None:
            return d
        nd = self.ndof
        return nd * st["perm"][d // nd] + d % nd

    # ---------------------------------------------------------------------- assemble (elements.py:165-200)
    def _bc_early(self, st):
        """Right-hand side and Dirichlet mask do not depend on K (elements.py:185-197 only writes brhs from the dof
        sets): in the reference's BC mode they are built on a side stream at the START of assemble() and the rhs
        travels to a pinned host buffer while the connectivity / coordinates are still uploading (PCIe is full
        duplex), so that brhs is on the host when assemble() returns, as in the reference."""
        import torch
        dev = st["conn"].device
        main = torch.cuda.current_stream(dev)
        side = side_streams(dev)[1]
        nrows = self.ndof * self.nnode
        if "rhs" in st:
            side.wait_stream(main)           # re-assembly: a solve on the main stream may still be reading rhs / dirmask
        if "dof_event" in st:
            side.wait_event(st["dof_event"])                     # dof sets were uploaded on the main stream
            for t in st["dof_sets"][2:]:
                t.record_stream(side)
        with torch.cuda.stream(side):
            if "rhs" not in st:
                st["rhs"] = torch.empty(nrows, dtype=torch.float64, device=dev)
                st["dirmask"] = torch.empty(nrows, dtype=torch.uint8, device=dev)
                st["rhs"].record_stream(main)
                st["dirmask"].record_stream(main)
            dir_d = self._solver_dofs(self.dofDir)
            neu_d = self._solver_dofs(self.dofNeu)
            torch.ops.feprogram.bc_rhs(dir_d, neu_d, 1.0, st["rhs"], st["dirmask"])   # elements.py:197 hardcoded 1
            st["bc_event"] = side.record_event()                 # the solver (main stream) waits for this one
            self._mark("rhs built", side)
            host = st.get("rhs_host")
            if host is None:
                host = st["rhs_host"] = torch.empty(nrows, dtype=torch.float64, pin_memory=True)
            host.copy_(st["rhs"], non_blocking=True)
            st["rhs_host_event"] = side.record_event()           # host readers (brhs, rhs_host) wait for this one
            self._mark("rhs on host", side)

    def rhs_host(self):
        """The assembled right-hand side as a host array (ndof * nnode,), reference dof numbering."""
        st = self._dev
        if st is None or not self._assembled:
            raise RuntimeError("assemble() must be called first")
        ev = st.get("rhs_host_event")
        if ev is not None and st["perm"] is None:
            ev.synchronize()
            return st["rhs_host"].numpy()
        rhs = st["rhs"]
        if st["perm"] is not None:
            rhs = rhs.view(-1, self.ndof)[st["perm"]].reshape(-1)
        return rhs.cpu().numpy()

    def assemble(self):
        import torch
        from . import ops
        st = self._state(defer_coords=True)
        et = _ELEM_CODES[self.elem.elemType]
        early_bc = getattr(self, "_bc_mode", "reference") == "reference" and st["perm"] is None
        st.pop("rhs_host_event", None)
        self._mark("state ready")
        if early_bc:
            self._bc_early(st)
        if "pattern" not in st:   # one-time symbolic phase for this mesh: CSR pattern, slot map, tile plan
            # overlaps the coordinate upload; the scipy-compatible row_ptr / col_idx (which neither the tile plan nor the
            # numeric kernel reads: they are a pure function of the node-level pattern) are expanded when K, the
            # reference-numbered system or the solver first asks for them (Pattern.row_ptr)
            # A lattice numbering's tiles depend on the connectivity alone: they are formed on the side stream while the
            # first half of the symbolic kernels runs on this one (the hook is called before the phase's host sync).
            csr_side = side_streams(st["conn"].device)[2]
            formed = {}

            def form_tiles(lattice):
                formed["tiles"] = ops.tile_formation(self.nelem, self.nnode, et, lattice, csr_side)

            # ... and the first half of their plan (ownership, node / halo lists, tile-ordered connectivity) needs the
            # incidence and the node degrees only: it is built on the side stream while fe_adjacency_fill runs here.
            def early_layout(eptr, einc, nptr):
                formed["layout"] = ops.plan_layout_early(st["conn_s"], et, self.nnode, eptr, einc, nptr,
                                                         formed.get("tiles"), csr_side)
            st["pattern"] = ops.symbolic(st["conn_s"], self.nnode, et, self.ndof, csr="lazy", between=form_tiles,
                                         after_fill=early_layout)
            self._mark("symbolic done")
            # the plan waits for the coordinate upload only if it needs coordinates (a lattice numbering does not)
            st["plan"] = ops.tile_plan(st["conn_s"], st["coords_s"], st["pattern"], coords_ready=self._coords_ready,
                                       formation=formed.get("tiles"), layout=formed.get("layout"))
            self._mark("plan done")
        self._coords_ready()
        pat = st["pattern"]
        dev = st["conn"].device
        if "vals" not in st:
            st["vals"] = torch.empty(pat.nnz, dtype=torch.float64, device=dev)
        ops.assemble_values(st["conn_s"], st["coords_s"], pat, st["plan"], st["vals"], self._operator)
        self._mark("numeric done")
        pat.csr_ready()                   # everything after assemble() may use row_ptr / col_idx on this stream
        st.pop("vals_ref", None)
        # boundary conditions: rhs + Dirichlet mask (matrix itself stays un-modified on the device)
        if early_bc:
            torch.cuda.current_stream(dev).wait_event(st["bc_event"])
        else:
            if "rhs" not in st:
                st["rhs"] = torch.empty(pat.nrows, dtype=torch.float64, device=dev)
                st["dirmask"] = torch.empty(pat.nrows, dtype=torch.uint8, device=dev)
            dir_d = self._solver_dofs(self.dofDir)
            neu_d = self._solver_dofs(self.dofNeu)
            if getattr(self, "_bc_mode", "reference") == "corrected":
                torch.ops.feprogram.bc_rhs(dir_d, neu_d[:0], 0.0, st["rhs"], st["dirmask"])
                edges = torch.from_numpy(self._load_edges).to(dev)
                if st["perm"] is not None:
                    edges = st["perm"][edges.to(torch.int64)].to(torch.int32).contiguous()
                torch.ops.feprogram.edge_load(edges, st["coords_s"], self._traction[0], self._traction[1], st["dirmask"],
                                              st["rhs"])
            else:
                torch.ops.feprogram.bc_rhs(dir_d, neu_d, 1.0, st["rhs"], st["dirmask"])   # elements.py:197 hardcoded 1
        self._K_host = None
        self._brhs_host = None
        self._assembled = True

    # --------------------------------------------------------------- lazily materialised scipy views
    def _ref_system(self):
        """(pattern, values, rhs, dirmask) in the REFERENCE numbering.  Without renumbering these are the solver's
        own arrays; with it the matrix is assembled on demand on the un-renumbered mesh (same element order and
        pair arithmetic => the same bits as a direct assembly) and rhs / mask are permuted back."""
        import torch
        from . import ops
        st = self._dev
        if st["perm"] is None:
            return st["pattern"], st["vals"], st["rhs"], st["dirmask"]
        if "pattern_ref" not in st:
            et = _ELEM_CODES[self.elem.elemType]
            st["pattern_ref"] = ops.symbolic(st["conn"], self.nnode, et, self.ndof)
            st["plan_ref"] = ops.tile_plan(st["conn"], st["coords"], st["pattern_ref"])
        pat = st["pattern_ref"]
        if "vals_ref" not in st:
            st["vals_ref"] = torch.empty(pat.nnz, dtype=torch.float64, device=st["conn"].device)
            ops.assemble_values(st["conn"], st["coords"], pat, st["plan_ref"], st["vals_ref"], self._operator)
        back = lambda v: v.view(-1, self.ndof)[st["perm"]].reshape(-1).contiguous()  # noqa: E731
        return pat, st["vals_ref"], back(st["rhs"]), back(st["dirmask"])

    def csr_prebc(self):
        """Pre-BC K as scipy CSR (indptr/indices bit-identical to the reference's fem.K.tocsr())."""
        pat, vals, _, _ = self._ref_system()
        return sp.csr_matrix((vals.cpu().numpy(), pat.col_idx.cpu().numpy(),
                              pat.row_ptr.cpu().numpy()), shape=(pat.nrows, pat.nrows))

    @property
    def K(self):
        if not self._assembled:
            n = self.nnode * self.ndof
            return sp.lil_matrix((n, n))
        if self._K_host is None:
            import torch
            pat, vals, rhs, dirmask = self._ref_system()
            vals = vals.clone()
            sym = getattr(self, "_bc_mode", "reference") == "corrected"      # zero Dirichlet values: rhs unchanged
            torch.ops.feprogram.apply_bc(1 if sym else 0, pat.row_ptr, pat.col_idx, vals,
                                         rhs.clone() if sym else rhs, dirmask)
            K = sp.csr_matrix((vals.cpu().numpy(), pat.col_idx.cpu().numpy(), pat.row_ptr.cpu().numpy()),
                              shape=(pat.nrows, pat.nrows))
            # Dirichlet rows keep only their unit diagonal (canonical form of the LIL deletes, elements.py:187-191)
            mask = dirmask.cpu().numpy().astype(bool)
            rows = np.repeat(np.arange(pat.nrows), np.diff(K.indptr))
            keep = ~mask[rows] | (K.indices == rows)
            counts = np.bincount(rows[keep], minlength=pat.nrows)
            K = sp.csr_matrix((K.data[keep], K.indices[keep], np.concatenate([[0], np.cumsum(counts)])),
                              shape=K.shape)
            self._K_host = K.tocsc()
        return self._K_host

    @K.setter
    def K(self, value):
        self._K_host = value

    @property
    def brhs(self):
        if not self._assembled:
            return sp.lil_matrix((self.nnode * self.ndof, 1))
        if self._brhs_host is None:
            self._brhs_host = sp.csc_matrix(np.array(self.rhs_host())[:, None])
        return self._brhs_host

    @brhs.setter
    def brhs(self, value):
        self._brhs_host = value

    # ------------------------------------------------------------------ solve (elements.py:241-243)
    def solveProblem(self):
        import torch
        st = self._dev
        if st is None or not self._assembled:
            raise RuntimeError("assemble() must be called before solveProblem()")
        pat = st["pattern"]
        x = torch.empty(pat.nrows, dtype=torch.float64, device=st["vals"].device)
        maxit = self.maxit if self.maxit is not None else 20 * pat.nrows + 100
        nodal = (pat.nptr, pat.nadj) if self.ndof == 2 else (None, None)     # node-block SpMV: 2 x 2 blocks only
        info = torch.ops.feprogram.pcg(pat.row_ptr, pat.col_idx, st["vals"], st["rhs"], st["dirmask"], x,
                                       self.rtol, 0.0, maxit, 50, *nodal)
        stop = float(info[3])
        self.solver_info = dict(iterations=int(info[0]), relres=float(info[1]), r0=float(info[2]),
                                converged=stop == 1.0, breakdown=stop < 0.0)
        if stop < 0.0 or not np.isfinite(self.solver_info["relres"]):
            # the reference's spsolve also handles indefinite systems (and warns on singular ones); CG cannot:
            # say so loudly instead of handing back a NaN / meaningless iterate as if it were the solution
            raise RuntimeError(
                "solveProblem: conjugate-gradient breakdown after %d iterations (relres %r): the assembled matrix is "
                "not positive definite on the free dofs (mixed element orientation, a degenerate element, or no "
                "Dirichlet support?)" % (self.solver_info["iterations"], self.solver_info["relres"]))
        if stop != 1.0:
            import warnings
            warnings.warn("solveProblem: PCG stopped at the iteration limit (%d) with relative residual %.3e > rtol %.1e"
                          % (self.solver_info["iterations"], self.solver_info["relres"], self.rtol), RuntimeWarning)
        st["x"] = x
        if st["perm"] is not None:
            x = x.view(-1, self.ndof)[st["perm"]].reshape(-1)  # back to the reference's dof numbering
        return x.cpu().numpy()
