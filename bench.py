#!/usr/bin/env python3
"""Benchmark of the FEM hot path on B200 (contract: one JSON line on stdout from rank 0).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg2|cfg3|cfg5|poisson] [--n NX]

Workload of the main line (BASELINE.json configs[1], "cfg2"): Quad4 structured mesh NX x NX (default 4096 ->
16.8 M elements), 2-dof reference operator, fp64.  For N > 1 every rank owns one NX x NX element block of an
NX x (N*NX) mesh (weak scaling): ghost-element assembly of its own rows, halo-exchange PCG.

A "step" = one pass of numeric assembly over the rank's elements (element matrices evaluated on the fly +
deterministic CSR reduction) + boundary-condition application.  `value` = elements assembled per second over
all ranks with inputs resident in HBM.  The symbolic phase is one-time per mesh and reported separately.  The
CG leg (SpMV GB/s, PCG iterations/s) is timed in the same run and reported under "cg"; `roofline` describes the
assembly kernel and `cg.roofline` the SpMV kernel.

`e2e` is the same metric through the public API (FemProblem ... assemble()) on a fresh problem with HOST
(pinned) connectivity / coordinates each step: host->device copies, tag conversion, symbolic phase, numeric
assembly, BCs, and the device->host read of the right-hand side + a checksum of K.

Additive keys: `parity_checked` (the timed kernel against the independent row kernel, bit for bit, at the timed
size), `strong_cfg4` (BASELINE configs[3], 8192 x 8192 elements split into N row blocks; the single-GPU time is
measured in the same run on rank 0), `dist_parity` (N > 1: a small problem solved to convergence over NCCL and
over peer memory against SciPy's direct solve of the assembled system), `e2e_full` (N = 1: host mesh ->
assemble -> solve -> U on the host at 512 x 512, beside the CPU port's assemble + spsolve).

`--config cfg3` (Quad9 2048^2, stiffness + mass) and `--config cfg5` (Tri3 5000^2 x 2, perturbed, shuffled
numbering) emit the same line for BASELINE configs[2] / configs[4] on one GPU (cfg5 also under torchrun on N GPUs:
general element-block partition of the shuffled mesh, "scaling": "strong"); `--config unstructured` a Delaunay-based
all-quad mesh; `--config poisson` (= `--operator
poisson`) for the scalar one-dof-per-node variant of cfg2 (FE_OP_POISSON = K[0::2, 0::2] of the reference operator).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "elements_assembled_per_s"
UNIT = "elements/s"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """DRAM bytes per launch from the committed ncu captures of exactly this workload (a pointer, not a live counter)."""
    for name in ("r02_dram_traffic.json", "r01_dram_traffic.json"):
        path = os.path.join(ROOT, "profiles", name)
        if os.path.exists(path):
            with open(path) as f:
                d = json.load(f)
            d["file"] = "profiles/" + name
            return d
    return {}


# ------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port on a bounded sample of the same workload
# ------------------------------------------------------------------------------------------------
def cpu_assembly_sample(sample_n: int):
    """Times the oracle's assembly (element matrices + COO->CSR duplicate summation, the reference's
    algorithm restated in vectorised NumPy/SciPy) on a sample_n x sample_n block of the workload."""
    from feprogram_b200 import mesh
    from oracle import fe_oracle as orc
    conn, coords, bc = mesh.structured_quad4(sample_n)
    t0 = time.perf_counter()
    K = orc.assemble_csr(conn - 1, coords)
    d, nn = orc.boundary_dofs(bc)
    orc.apply_bc_reference(K, d, nn)
    dt = time.perf_counter() - t0
    return conn.shape[0] / dt, dt


def _cpu_worker(sample_n):
    return cpu_assembly_sample(sample_n)[0]


def cpu_assembly_pool(sample_n: int, procs: int):
    """The oracle port on every host core: `procs` independent sample_n x sample_n blocks of the workload in a
    process pool (the port itself is single-threaded NumPy/SciPy, like the reference).  Assembly cost is linear
    in the element count, so independent blocks measure the same per-element rate as one large mesh.
    Returns (elements/s over all processes, wall seconds)."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    with ctx.Pool(procs) as pool:
        pool.map(_cpu_worker, [32] * procs)                 # start the workers, import numpy/scipy
        t0 = time.perf_counter()
        pool.map(_cpu_worker, [sample_n] * procs)
        dt = time.perf_counter() - t0
    return procs * sample_n * sample_n / dt, dt


def cpu_full_path(n: int):
    """The CPU port's whole path on an n x n Quad4 mesh: assemble + BCs + SciPy's direct solve (what the reference's
    assemble() + solveProblem() do, elements.py:165-200, 241-243).  Returns seconds (assemble, solve)."""
    from feprogram_b200 import mesh
    from oracle import fe_oracle as orc
    conn, coords, bc = mesh.structured_quad4(n)
    t0 = time.perf_counter()
    _, _, _, U = orc.solve_reference_path(conn - 1, coords, bc)
    return time.perf_counter() - t0, U


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path.  The reference is Python and
    cannot travel to the GPU box, so the arm times the oracle port (oracle/fe_oracle.py), which restates
    elements.py:153-200 with vectorised NumPy + SciPy's COO->CSR (single host thread, like the reference)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample_n = min(args.cpu_sample, 512)
    procs = os.cpu_count() or 1
    vals, times = [], []
    for _ in range(max(1, args.steps)):
        v, dt = cpu_assembly_pool(sample_n, procs)
        vals.append(v)
        times.append(dt)
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "quad4_%dx%d_2dof_reference_operator" % (args.n, args.n),
                   "sample": "%d blocks of %dx%d elements per step, one per host core (assembly is linear in the element "
                             "count: the per-element rate of the blocks is that of the full mesh)" % (procs, sample_n, sample_n)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "host_cores": os.cpu_count(), "kind": "port",
                         "sample": "oracle/fe_oracle.py assemble_csr + apply_bc_reference on %d independent %dx%d Quad4 "
                                   "blocks (%d elements) per step, one process per host core"
                                   % (procs, sample_n, sample_n, procs * sample_n * sample_n)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# helpers of our arm
# ------------------------------------------------------------------------------------------------
def _timed(torch, fn, reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def bc_dofs(bc):
    return [np.unique(np.concatenate([2 * (np.array(sorted(s), dtype=np.int64) - 1),
                                      2 * (np.array(sorted(s), dtype=np.int64) - 1) + 1])) for s in (bc[0], bc[3])]


def single_gpu_case(torch, dev, kind, n, reps, cg_iters, perturb=False, shuffle=False, ops_list=(0,), ndof=2,
                    unstructured=False):
    """Resident-input timing of one mesh on one GPU: symbolic phase, numeric assembly per operator, SpMV, PCG.
    Returns a dict; used for cfg3 / cfg4 / cfg5 and the strong-scaling single-GPU reference."""
    from feprogram_b200 import mesh, ops
    if unstructured:       # Delaunay-based, random numbering: `shuffle` then only asks for the renumbered solver copy
        conn_h, coords_h, bc = (mesh.unstructured_quad4 if kind == 4 else mesh.unstructured_tri3)(n, seed=3, dtype=np.int32)
    else:
        gen = {4: mesh.structured_quad4, 9: mesh.structured_quad9, 3: mesh.structured_tri3}[kind]
        conn_h, coords_h, bc = gen(n, dtype=np.int32)
        mx = 2 * n if kind == 9 else n
        if perturb:
            coords_h = mesh.perturb_interior(coords_h, mx, mx, 0.2, seed=7)
        if shuffle:
            conn_h, coords_h, bc, _ = mesh.shuffle_numbering(conn_h, coords_h, bc, seed=11)
    conn = torch.from_numpy(conn_h - 1).to(dev)
    coords = torch.from_numpy(coords_h).to(dev)
    nnode = coords_h.shape[0]
    renumbered = False
    if shuffle:      # what FemProblem does for numberings without locality: solve on a Z-curve renumbered copy
        perm, order = ops.locality_order(coords)
        conn = perm[conn.to(torch.int64)].to(torch.int32).contiguous()
        coords = coords[order].contiguous()
        renumbered = True
    # warm the symbolic / plan path on a small mesh of the same kind (module load, first-use allocations), then time it
    wc, wx, _ = (mesh.structured_quad4 if kind == 4 else mesh.structured_quad9 if kind == 9 else mesh.structured_tri3)(48, dtype=np.int32)
    wconn, wxy = torch.from_numpy(wc - 1).to(dev), torch.from_numpy(wx).to(dev)
    ops.tile_plan(wconn, wxy, ops.symbolic(wconn, wx.shape[0], kind, ndof))
    del wconn, wxy
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pat = ops.symbolic(conn, nnode, kind, ndof)
    plan = ops.tile_plan(conn, coords, pat)
    torch.cuda.synchronize()
    sym_s = time.perf_counter() - t0
    vals = torch.empty(pat.nnz, dtype=torch.float64, device=dev)
    asm_bytes = conn.shape[0] * kind * 4 + nnode * 16 + pat.nnz * 8
    out = {"elements": int(conn.shape[0]), "nnz": int(pat.nnz), "symbolic_phase_s": sym_s, "renumbered": renumbered,
           "plan": None if plan is None else {k: plan.stats.get(k) for k in ("mode", "ntiles", "kernel", "max_nodes", "max_halo")},
           "kernel": ("k_assemble_ws" if plan is not None and plan.warp_specialised else
                      "k_assemble_tiled" if plan is not None else "k_assemble_rows"),
           "algorithmic_bytes_per_launch": asm_bytes, "assembly": {}}
    for op in ops_list:
        for _ in range(3):
            ops.assemble_values(conn, coords, pat, plan, vals, op)
        ms = _timed(torch, lambda: ops.assemble_values(conn, coords, pat, plan, vals, op), reps)
        out["assembly"][{0: "stiffness", 1: "weighted", 2: "mass", 3: "stiffness"}[op]] = {
            "ms_per_launch": ms, "elements_per_s": conn.shape[0] / ms * 1e3, "gbs": asm_bytes / ms / 1e6}
    ops.assemble_values(conn, coords, pat, plan, vals, ops_list[0] if ndof == 1 else 0)
    if plan is not None:      # the timed kernel against the independent row kernel, bit for bit, at this size
        v_r = torch.full((pat.nnz,), float("nan"), dtype=torch.float64, device=dev)
        ops.assemble_values(conn, coords, pat, None, v_r, ops_list[0])
        out["parity_checked"] = {"kernel": out["kernel"], "against": "k_assemble_rows (independent kernel), NaN-prefilled output",
                                 "tiled_eq_rows": bool(torch.equal(vals, v_r)), "entries": int(pat.nnz)}
        del v_r
    if cg_iters > 0:
        d_dofs, n_dofs = bc_dofs(bc)
        if ndof == 1:
            d_dofs, n_dofs = np.unique(d_dofs // 2), np.unique(n_dofs // 2)
        if renumbered and ndof == 1:
            d_dofs = perm[torch.from_numpy(d_dofs).to(dev)].cpu().numpy()
            n_dofs = perm[torch.from_numpy(n_dofs).to(dev)].cpu().numpy()
        elif renumbered:
            d_dofs = (2 * perm[torch.from_numpy(d_dofs // 2).to(dev)] + torch.from_numpy(d_dofs % 2).to(dev)).cpu().numpy()
            n_dofs = (2 * perm[torch.from_numpy(n_dofs // 2).to(dev)] + torch.from_numpy(n_dofs % 2).to(dev)).cpu().numpy()
        rhs = torch.empty(pat.nrows, dtype=torch.float64, device=dev)
        dirmask = torch.empty(pat.nrows, dtype=torch.uint8, device=dev)
        torch.ops.feprogram.bc_rhs(torch.from_numpy(d_dofs).to(dev), torch.from_numpy(n_dofs).to(dev), 1.0, rhs, dirmask)
        x = torch.randn(pat.nrows, dtype=torch.float64, device=dev)
        y = torch.empty_like(x)
        for _ in range(3):
            torch.ops.feprogram.spmv(pat.row_ptr, pat.col_idx, vals, x, y, None)
        ms_spmv = _timed(torch, lambda: torch.ops.feprogram.spmv(pat.row_ptr, pat.col_idx, vals, x, y, None), max(reps, 5))
        sol = torch.empty(pat.nrows, dtype=torch.float64, device=dev)
        nodal = (pat.nptr, pat.nadj) if ndof == 2 else (None, None)
        torch.ops.feprogram.pcg(pat.row_ptr, pat.col_idx, vals, rhs, dirmask, sol, 0.0, 0.0, 10, 10, *nodal)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        info = torch.ops.feprogram.pcg(pat.row_ptr, pat.col_idx, vals, rhs, dirmask, sol, 0.0, 0.0, cg_iters, cg_iters,
                                       *nodal)
        b.record()
        torch.cuda.synchronize()
        spmv_bytes = 12 * pat.nnz + 20 * pat.nrows
        out["cg"] = {"spmv_ms": ms_spmv, "spmv_gbs": spmv_bytes / ms_spmv / 1e6, "spmv_bytes": spmv_bytes,
                     "pcg_iters_per_s": int(info[0]) / (a.elapsed_time(b) * 1e-3)}
    return out


def dist_parity_check(torch, dist, fdist, rank, world, dev, comm):
    """N > 1: a 64 x (64 N) problem solved to convergence over NCCL and over peer memory, compared on rank 0 with
    SciPy's direct solve (the reference's solver call, elements.py:241-243) of the system this library assembled."""
    from feprogram_b200 import mesh
    from feprogram_b200.elements import FemProblem
    m, bc = fdist.structured_quad4_block(64, 64, rank, world)
    p = fdist.DistFemProblem(m, dev, 4, comm)
    p.assemble()
    p.set_bc(bc["dir"], bc["neu"])
    x_nccl = p.solve(rtol=1e-11, maxit=20000, check_every=50, p2p=False)
    info_nccl = dict(p.info)
    res = {"iters": info_nccl["iterations"], "converged": bool(info_nccl["converged"])}
    x_p2p = None
    if p.enable_p2p():
        x_p2p = p.solve(rtol=1e-11, maxit=20000, check_every=50, p2p=True)
        res["p2p_iters"] = p.info["iterations"]
        res["p2p_converged"] = bool(p.info["converged"])
        diff = (x_p2p[: p.nrows_owned] - x_nccl[: p.nrows_owned]).abs().max()
        ref = x_nccl[: p.nrows_owned].abs().max()
        t = torch.stack([diff, ref])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res["nccl_vs_p2p_max_rel"] = float(t[0] / t[1])
        res["nccl_eq_p2p"] = bool(t[0].item() == 0.0)          # bit-identical at N = 2; reduction order differs beyond
    own = [None] * world
    dist.all_gather_object(own, (m.node_gid[: m.n_owned], x_nccl[: p.nrows_owned].cpu().numpy(),
                                 None if x_p2p is None else x_p2p[: p.nrows_owned].cpu().numpy()))
    if rank == 0:
        import scipy.sparse.linalg as spla
        conn, coords, bcs = mesh.structured_quad4(64, 64 * world)
        fem = FemProblem(conn, coords, device=dev)
        fem.setMatrix()
        fem.setBoundaryConditions(bcs, verbose=False)
        fem.assemble()
        U = spla.spsolve(fem.K, np.asarray(fem.brhs.todense()).ravel())
        nn = coords.shape[0]
        for key, col in (("err", 1), ("err_p2p", 2)):
            if own[0][col] is None:
                continue
            Ud = np.zeros(2 * nn)
            for part in own:
                gid = np.asarray(part[0])
                Ud[2 * gid] = part[col][0::2]
                Ud[2 * gid + 1] = part[col][1::2]
            res[key] = float(np.linalg.norm(Ud - U) / np.linalg.norm(U))
        res["against"] = "scipy.sparse.linalg.spsolve of this library's K, brhs on the 64 x %d mesh" % (64 * world)
    del p
    return res


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(torch, local):
    """N > 1: run this rank's host threads on the NUMA node its GPU hangs off, BEFORE any pinned buffer is allocated
    (first-touch places the pages there), so that eight simultaneous uploads do not all cross one socket link.
    Best effort: returns the node, or None when the topology files are not there."""
    try:
        pr = torch.cuda.get_device_properties(local)
        bus = "%04x:%02x:%02x.0" % (int(pr.pci_domain_id), int(pr.pci_bus_id), int(pr.pci_device_id))
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:
        return None
    return None


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    numa_node = bind_to_gpu_numa_node(torch, local) if world > 1 else None
    # the contract is ONE JSON line on stdout: anything libraries print there (NCCL's version banner) goes to stderr
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from feprogram_b200 import build
    if rank == 0:
        build.build()
    if world > 1:
        dist.barrier()
    from feprogram_b200 import mesh, ops
    from feprogram_b200.elements import FemProblem

    hbm_peak, peak_src = peaks()
    n = args.n
    K, W = args.steps, args.warmup

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.config != "cfg2":
        return run_other_config(args, torch, dev, hbm_peak, peak_src, real_stdout)

    # ---- inputs resident in HBM -------------------------------------------------------------------
    parity = None
    if world == 1:
        conn_h, coords_h, bc = mesh.structured_quad4(n, dtype=np.int32)
        part = None
        conn = torch.from_numpy(conn_h - 1).to(dev)
        coords = torch.from_numpy(coords_h).to(dev)
        nnode = coords_h.shape[0]
        mkplan = (lambda p: None) if args.kernel == "rows" else (lambda p: ops.tile_plan(conn, coords, p, kernel=args.kernel))
        pat = ops.symbolic(conn, nnode, 4, 2)
        plan = mkplan(pat)
        torch.cuda.synchronize()
        sym_times = []                          # rebuilds: exclude CUDA context / module load / first cudaMallocs
        for _ in range(3):
            del pat, plan
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            pat = ops.symbolic(conn, nnode, 4, 2)
            plan = mkplan(pat)
            torch.cuda.synchronize()
            sym_times.append(time.perf_counter() - t0)
        symbolic_s = sorted(sym_times)[1]
        nelem_local = conn.shape[0]
        d_dofs, n_dofs = bc_dofs(bc)
        dir_d = torch.from_numpy(d_dofs).to(dev)
        neu_d = torch.from_numpy(n_dofs).to(dev)
        vals = torch.empty(pat.nnz, dtype=torch.float64, device=dev)
        rhs = torch.empty(pat.nrows, dtype=torch.float64, device=dev)
        dirmask = torch.empty(pat.nrows, dtype=torch.uint8, device=dev)

        def step():
            ops.assemble_values(conn, coords, pat, plan, vals, 0)
            torch.ops.feprogram.bc_rhs(dir_d, neu_d, 1.0, rhs, dirmask)
        launches_per_step = 3  # k_assemble + k_mark_dirichlet + k_set_load (2 memsets not counted)
        asm_bytes = nelem_local * 4 * 4 + nnode * 16 + pat.nnz * 8
        spmv_bytes = 12 * pat.nnz + 20 * pat.nrows
        kernel_name = ("k_assemble_ws" if plan is not None and plan.warp_specialised else
                       "k_assemble_tiled" if plan is not None else "k_assemble_rows")
        # ---- parity of the timed kernel at the timed size (outside every timed region) ----
        if plan is not None:
            v_t = torch.full((pat.nnz,), float("nan"), dtype=torch.float64, device=dev)
            ops.assemble_values(conn, coords, pat, plan, v_t, 0)
            v_r = torch.empty(pat.nnz, dtype=torch.float64, device=dev)
            ops.assemble_values(conn, coords, pat, None, v_r, 0)
            parity = {"kernel": kernel_name, "against": "k_assemble_rows (independent kernel), NaN-prefilled output",
                      "tiled_eq_rows": bool(torch.equal(v_t, v_r)), "nan_free": not bool(torch.isnan(v_t).any()),
                      "entries": int(pat.nnz)}
            del v_t, v_r
    else:
        from feprogram_b200 import dist as fdist
        part = fdist.build_structured_block(n, rank, world, dev, args.ny_per_rank or n, args.halo)
        symbolic_s = part.symbolic_s
        nelem_local = part.nelem_owned
        step = part.assemble_step
        launches_per_step = part.launches_per_step
        asm_bytes = part.asm_bytes
        spmv_bytes = part.spmv_bytes
        kernel_name = ("k_assemble_ws" if part.p.plan is not None and part.p.plan.warp_specialised else
                       "k_assemble_tiled" if part.tiled else "k_assemble_rows")

    # ---- assembly: W warm-ups, K timed steps ------------------------------------------------------
    clk = ClockSampler(local)
    clk.__enter__()                       # samples clocks from the warm-up to the end of the CG leg
    time.sleep(0.4)
    for _ in range(W):
        step()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(K):
        step()
    ev1.record()
    barrier()
    ms_step = ev0.elapsed_time(ev1) / K
    # kernel-only timing of the dominant kernel (same stream, CUDA events)
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    asm_only = part.assemble_kernel_only if part is not None else (
        lambda: ops.assemble_values(conn, coords, pat, plan, vals, 0))
    barrier()
    k0.record()
    for _ in range(K):
        asm_only()
    k1.record()
    barrier()
    ms_kernel = k0.elapsed_time(k1) / K

    # ---- CG leg: SpMV bandwidth and PCG iteration rate ------------------------------------------------
    cg_iters = args.cg_iters
    if part is None:
        x = torch.randn(pat.nrows, dtype=torch.float64, device=dev)
        y = torch.empty_like(x)
        for _ in range(W):
            torch.ops.feprogram.spmv(pat.row_ptr, pat.col_idx, vals, x, y, None)
        barrier()
        ms_spmv = _timed(torch, lambda: torch.ops.feprogram.spmv(pat.row_ptr, pat.col_idx, vals, x, y, None), max(K, 5))
        sol = torch.empty(pat.nrows, dtype=torch.float64, device=dev)
        torch.ops.feprogram.pcg(pat.row_ptr, pat.col_idx, vals, rhs, dirmask, sol, 0.0, 0.0, 10, 10, pat.nptr, pat.nadj)
        barrier()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        info = torch.ops.feprogram.pcg(pat.row_ptr, pat.col_idx, vals, rhs, dirmask, sol, 0.0, 0.0, cg_iters, cg_iters,
                                       pat.nptr, pat.nadj)
        p1.record()
        barrier()
        ms_pcg = p0.elapsed_time(p1)
        pcg_done = int(info[0])
        del x, y, sol
    else:
        ms_spmv = part.time_spmv(W, max(K, 5))
        ms_pcg, pcg_done = part.time_pcg(cg_iters)

    clk.__exit__()
    # ---- max over ranks -------------------------------------------------------------------------------
    t = torch.tensor([ms_step, ms_kernel, ms_spmv, ms_pcg, symbolic_s], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(nelem_local), float(asm_bytes), float(spmv_bytes)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_step, ms_kernel, ms_spmv, ms_pcg, symbolic_s = t.tolist()
    nelem_total, asm_bytes_total, spmv_bytes_total = tot.tolist()

    # ---- e2e through the public API with host buffers ---------------------------------------------------
    e2e = None
    e2e_full = None
    if world == 1:
        conn1 = torch.from_numpy(conn_h).pin_memory()          # 1-based tags, host pinned
        xy_p = torch.from_numpy(coords_h).pin_memory()
        conn1_np, xy_np = conn1.numpy(), xy_p.numpy()

        def e2e_step():
            fem = FemProblem(conn1_np, xy_np, device=dev)
            fem.setMatrix()
            fem.dofDir, fem.dofNeu = d_dofs, n_dofs            # what setBoundaryConditions computes (host sets)
            fem.assemble()
            chk = fem._dev["vals"].sum().item()                 # forces completion of the assembly; 8-byte D2H
            b = fem.rhs_host()                                  # pinned host copy of the right-hand side (D2H inside assemble)
            return chk, float(b[-1])
        reps = max(1, min(K, 3))
        e2e_step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            chk, _ = e2e_step()
        torch.cuda.synchronize()
        e2e_s = (time.perf_counter() - t0) / reps
        e2e = {"value": nelem_total / e2e_s, "unit": UNIT,
               "h2d_bytes_per_step": int(conn_h.nbytes + coords_h.nbytes + d_dofs.nbytes + n_dofs.nbytes),
               "d2h_bytes_per_step": int(pat.nrows * 8 + 8),
               "ms_per_step": 1e3 * e2e_s, "includes": "H2D, tag conversion, symbolic phase (node-level CSR pattern; the dof-level "
               "row_ptr / col_idx are a pure function of it and are expanded when K or the solver first reads them), "
               "tile plan, numeric assembly, BCs, D2H rhs + checksum of all of K",
               "checksum": chk}
        del conn1, xy_p
    elif part is not None:
        e2e = part.e2e(K)

    # ---- strong scaling on BASELINE configs[3] (8192 x 8192 elements over N row blocks) -------------------
    strong = None
    if not args.no_strong:
        strong = strong_cfg4(args, torch, dist if world > 1 else None, rank, world, dev, part)

    # ---- N > 1: converged-solve parity of the distributed PCG (NCCL and peer memory) ------------------------
    dist_parity = None
    if world > 1:
        from feprogram_b200 import dist as fdist
        dist_parity = dist_parity_check(torch, dist, fdist, rank, world, dev, part.comm)

    # ---- CPU baseline + full path (rank 0, N == 1 only) -------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        procs = os.cpu_count() or 1
        sn = min(args.cpu_sample, 512)
        v, dt = cpu_assembly_pool(sn, procs)
        cpu = {"value": v, "unit": UNIT, "cores": procs, "host_cores": os.cpu_count(), "kind": "port",
               "sample": "oracle/fe_oracle.py assemble_csr + apply_bc_reference on %d independent %dx%d Quad4 blocks "
                         "(%d elements, %.1f s wall), one process per host core" % (procs, sn, sn, procs * sn * sn, dt)}
        # full path, one repetition: host mesh -> assemble -> solve -> U on the host, ours beside the port's
        nf = args.full_n
        cf, xf, bf = mesh.structured_quad4(nf, dtype=np.int32)
        for rep in range(2):                  # first repetition warms the allocator
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fem = FemProblem(cf, xf, device=dev)
            fem.setMatrix()
            fem.setBoundaryConditions(bf, verbose=False)
            fem.assemble()
            t1 = time.perf_counter()
            U = fem.solveProblem()
            gpu_s = time.perf_counter() - t0
            solve_s = time.perf_counter() - t1
        cpu_s, U_cpu = cpu_full_path(nf)
        e2e_full = {"workload": "quad4_%dx%d: FemProblem(...) -> setMatrix -> setBoundaryConditions -> assemble -> solveProblem, "
                                "host arrays in, U on the host out" % (nf, nf),
                    "ours_s": gpu_s, "ours_solve_s": solve_s, "pcg_iterations": fem.solver_info["iterations"],
                    "relres": fem.solver_info["relres"],
                    "cpu_port_s": cpu_s, "cpu_port": "oracle assemble + BCs + scipy spsolve, 1 core",
                    "speedup": cpu_s / gpu_s, "rel_l2_vs_cpu_port": float(np.linalg.norm(U - U_cpu) / np.linalg.norm(U_cpu))}

    traffic = ncu_traffic() if (world == 1 and n == 4096) else {}
    if rank == 0:
        value = nelem_total / (ms_step * 1e-3)
        asm_gbs = (asm_bytes_total / world) / (ms_kernel * 1e-3) / 1e9
        spmv_gbs = (spmv_bytes_total / world) / (ms_spmv * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "quad4_%dx%d_2dof_reference_operator%s" % (n, (args.ny_per_rank or n) * world if world > 1 else n,
                                                                              "" if world == 1 else "_row_blocks"),
                       "elements": int(nelem_total), "l2": "inputs exceed L2 (no flush needed)" if nelem_total / world > 2e6 else "inputs fit L2",
                       "symbolic_phase_s": symbolic_s, "partition": "dp%d element row blocks, owned-row assembly with ghost elements" % world,
                       "host_numa_node_rank0": numa_node},
            "roofline": {"kernel": kernel_name, "bound": "hbm", "achieved": asm_gbs, "peak": hbm_peak,
                         "unit": "GB/s", "frac": asm_gbs / hbm_peak,
                         "traffic": traffic.get(kernel_name),
                         "traffic_source": ("ncu --set full capture of this workload, %s (not a live counter)" % traffic.get("file")
                                            if traffic.get(kernel_name) else None),
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": asm_bytes_total / world, "ms_per_launch": ms_kernel},
            "cg": {"spmv_ms": ms_spmv, "spmv_gbs": spmv_gbs, "spmv_frac": spmv_gbs / hbm_peak,
                   "pcg_iters_timed": pcg_done, "halo": (part.halo if part is not None else None), "pcg_iters_per_s": pcg_done / (ms_pcg * 1e-3) if ms_pcg > 0 else None,
                   "roofline": {"kernel": "k_spmv", "bound": "hbm", "achieved": spmv_gbs, "peak": hbm_peak, "unit": "GB/s",
                                "frac": spmv_gbs / hbm_peak, "traffic": traffic.get("k_spmv"),
                                "traffic_source": ("ncu --set full capture, %s (not a live counter)" % traffic.get("file")
                                                   if traffic.get("k_spmv") else None),
                                "algorithmic_bytes_per_launch": spmv_bytes_total / world}},
            "parity_checked": parity, "dist_parity": dist_parity, "strong_cfg4": strong,
            "cpu_baseline": cpu, "e2e": e2e, "e2e_full": e2e_full, "gpu_launches": launches_per_step * K,
            "clocks": clk.summary(),
        }
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    if world > 1:
        dist.destroy_process_group()


def strong_cfg4(args, torch, dist, rank, world, dev, part):
    """BASELINE configs[3]: Quad4 8192 x 8192 (67 M elements).  N = 1: the whole mesh on this GPU.  N > 1: N row
    blocks of 8192 x 8192/N elements (assembly step and PCG over all ranks, max over ranks) AND, on rank 0 alone,
    the whole mesh on one GPU in the same run -- so speedup_vs_n1 divides two numbers of the same box and build."""
    nx = args.strong_n
    reps = max(3, min(args.steps, 5))
    cg = min(args.cg_iters, 100)
    out = {"workload": "quad4_%dx%d_2dof_reference_operator" % (nx, nx), "elements": nx * nx}
    if world == 1:
        r = single_gpu_case(torch, dev, 4, nx, reps, cg)
        out.update({"ms_per_step": r["assembly"]["stiffness"]["ms_per_launch"],
                    "elements_per_s": r["assembly"]["stiffness"]["elements_per_s"],
                    "assembly_gbs": r["assembly"]["stiffness"]["gbs"], "spmv_gbs": r["cg"]["spmv_gbs"],
                    "pcg_iters_per_s": r["cg"]["pcg_iters_per_s"], "speedup_vs_n1": 1.0, "kernel": r["kernel"]})
        return out
    from feprogram_b200 import dist as fdist
    torch.cuda.synchronize()
    p = fdist.build_structured_block(nx, rank, world, dev, nx // world, args.halo, comm=part.comm)
    for _ in range(3):
        p.assemble_step()
    dist.barrier()
    torch.cuda.synchronize()
    ms = _timed(torch, p.assemble_step, reps)
    ms_pcg, iters = p.time_pcg(cg)
    t = torch.tensor([ms, ms_pcg], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_pcg = t.tolist()
    halo = p.halo
    del p
    torch.cuda.empty_cache()
    single = None
    if rank == 0:
        r = single_gpu_case(torch, dev, 4, nx, reps, cg)
        single = {"ms_per_step": r["assembly"]["stiffness"]["ms_per_launch"], "pcg_iters_per_s": r["cg"]["pcg_iters_per_s"]}
    dist.barrier()
    if rank == 0:
        out.update({"ms_per_step": ms, "elements_per_s": nx * nx / ms * 1e3, "pcg_iters_per_s": iters / (ms_pcg * 1e-3),
                    "halo": halo, "n1_same_run": single,
                    "speedup_vs_n1": {"assembly": single["ms_per_step"] / ms,
                                      "pcg": iters / (ms_pcg * 1e-3) / single["pcg_iters_per_s"]}})
    return out


def cfg5_multi_gpu(args, torch, dev, hbm_peak, peak_src, real_stdout):
    """--config cfg5 with N > 1 (BASELINE configs[4]): perturbed Tri3 mesh, randomly shuffled node numbering, general
    element-block partition over the ranks (partition_mesh: Z-curve local numbering), owned-row assembly without a
    collective, peer-memory / NCCL PCG.  Every rank generates the same synthetic mesh on the host and keeps its share."""
    import torch.distributed as dist
    from feprogram_b200 import dist as fdist
    from feprogram_b200 import mesh
    rank, world = dist.get_rank(), dist.get_world_size()
    n = args.n if args.n != 4096 else 5000
    K, W = args.steps, max(3, args.warmup)
    conn, coords, bc = mesh.structured_tri3(n, dtype=np.int32)
    coords = mesh.perturb_interior(coords, n, n, 0.2, 0)
    conn, coords, bc, _ = mesh.shuffle_numbering(conn, coords, bc, 1)
    m = fdist.exchange_send_lists(fdist.partition_mesh(conn - 1, coords, rank, world, locality=True))
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
    symbolic_s = p.symbolic_s
    p.set_bc(dir_l, neu_l)
    halo = "peer-memory (NVLink loads + flag-gated scalar exchange)" if (args.halo == "p2p" and p.enable_p2p()) else "nccl"
    clk = ClockSampler(dev.index or 0)
    clk.__enter__()

    def timed(f, reps, warm):
        for _ in range(warm):
            f()
        torch.cuda.synchronize()
        dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            f()
        b.record()
        dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / reps], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def step():
        p.assemble()
        p.apply_bc()
    ms_step = timed(step, K, W)
    ms_asm = timed(p.assemble, K, 1)
    x = torch.randn(p.nrows_local, dtype=torch.float64, device=dev)
    y = torch.empty_like(x)
    ms_spmv = timed(lambda: p.spmv_owned(x, y), max(K, 5), 3)
    iters = args.cg_iters
    p.solve(rtol=0.0, maxit=10, check_every=10, p2p=halo != "nccl")
    ms_pcg = timed(lambda: p.solve(rtol=0.0, maxit=iters, check_every=iters, p2p=halo != "nccl"), 1, 0)
    clk.__exit__()
    stats = torch.tensor([p.pat.nnz, m.conn.shape[0], m.n_owned, m.nnode - m.n_owned, symbolic_s], dtype=torch.float64, device=dev)
    allst = [torch.empty_like(stats) for _ in range(world)]
    dist.all_gather(allst, stats)
    if rank == 0:
        nnz = sum(float(s[0]) for s in allst)
        asm_bytes = (nelem_total * 3 * 4 + nnode_total * 16 + nnz * 8) / world
        spmv_bytes = (12 * nnz + 20 * 2 * nnode_total) / world
        asm_gbs, spmv_gbs = asm_bytes / ms_asm / 1e6, spmv_bytes / ms_spmv / 1e6
        line = {"metric": METRIC, "value": nelem_total / ms_step * 1e3, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": "tri3_%dx%dx2_perturbed_shuffled_numbering" % (n, n), "elements": nelem_total,
                           "l2": "inputs exceed L2 (no flush needed)", "symbolic_phase_s": max(float(s[4]) for s in allst),
                           "partition": "dp%d element blocks of the shuffled mesh (general partition, Z-curve local numbering), "
                                        "owned-row assembly with ghost elements" % world,
                           "ghost_nodes_per_rank": [int(s[3]) for s in allst],
                           "tile_plan": None if p.plan is None else p.plan.stats.get("mode")},
                "roofline": {"kernel": "k_assemble_tiled" if p.plan is not None else "k_assemble_rows", "bound": "hbm",
                             "achieved": asm_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": asm_gbs / hbm_peak, "traffic": None,
                             "peak_source": peak_src, "algorithmic_bytes_per_launch": asm_bytes, "ms_per_launch": ms_asm,
                             "per": "GPU (owned elements, nodes and non-zeros of a rank)"},
                "cg": {"spmv_ms": ms_spmv, "spmv_gbs": spmv_gbs, "spmv_frac": spmv_gbs / hbm_peak, "pcg_iters_timed": iters,
                       "pcg_iters_per_s": iters / ms_pcg * 1e3, "halo": halo},
                "cpu_baseline": None, "e2e": None, "gpu_launches": 3 * K, "clocks": clk.summary()}
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    dist.barrier()
    dist.destroy_process_group()


def run_other_config(args, torch, dev, hbm_peak, peak_src, real_stdout):
    """--config cfg3 (Quad9 2048^2, stiffness + mass) / cfg5 (Tri3 5000^2 x 2, perturbed, shuffled tags) / poisson (the
    scalar one-dof-per-node variant of cfg2, SURVEY §8d: 104 algorithmic bytes per element): one GPU."""
    K, W = args.steps, max(3, args.warmup)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        if args.config != "cfg5":
            raise SystemExit("--config %s is a one-GPU line; only cfg2 (default) and cfg5 run on N > 1" % args.config)
        return cfg5_multi_gpu(args, torch, dev, hbm_peak, peak_src, real_stdout)
    clk = ClockSampler(dev.index or 0)
    clk.__enter__()
    if args.config == "cfg3":
        n = args.n if args.n != 4096 else 2048
        r = single_gpu_case(torch, dev, 9, n, K, args.cg_iters, ops_list=(0, 2))
        workload = "quad9_%dx%d_2dof_stiffness_and_mass" % (n, n)
    elif args.config == "unstructured":
        n = args.n if args.n != 4096 else 800
        r = single_gpu_case(torch, dev, 4, n, K, args.cg_iters, shuffle=True, unstructured=True)
        workload = "quad4_unstructured_delaunay_split_%d_points_per_side_random_numbering" % (n + 1)
    elif args.config == "poisson":
        n = args.n
        r = single_gpu_case(torch, dev, 4, n, K, args.cg_iters, ops_list=(3,), ndof=1)
        workload = "quad4_%dx%d_1dof_scalar_poisson_operator" % (n, n)
    else:
        n = args.n if args.n != 4096 else 5000
        r = single_gpu_case(torch, dev, 3, n, K, args.cg_iters, perturb=True, shuffle=True)
        workload = "tri3_%dx%dx2_perturbed_shuffled_numbering" % (n, n)
    clk.__exit__()
    a = r["assembly"]["stiffness"]
    line = {"metric": METRIC, "value": a["elements_per_s"], "unit": UNIT, "n_gpus": 1, "steps": K, "warmup": W,
            "ms_per_step": a["ms_per_launch"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "elements": r["elements"], "l2": "inputs exceed L2 (no flush needed)",
                       "symbolic_phase_s": r["symbolic_phase_s"], "renumbered": r["renumbered"], "plan": r.get("plan")},
            "roofline": {"kernel": r["kernel"], "bound": "hbm", "achieved": a["gbs"], "peak": hbm_peak, "unit": "GB/s",
                         "frac": a["gbs"] / hbm_peak, "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": r["algorithmic_bytes_per_launch"], "ms_per_launch": a["ms_per_launch"]},
            "assembly": {k: dict(v, frac=v["gbs"] / hbm_peak) for k, v in r["assembly"].items()},
            "cg": dict(r.get("cg", {}), spmv_frac=r["cg"]["spmv_gbs"] / hbm_peak) if "cg" in r else None,
            "parity_checked": r.get("parity_checked"),
            "cpu_baseline": None, "e2e": None, "gpu_launches": K, "clocks": clk.summary()}
    real_stdout.write(json.dumps(line) + "\n")
    real_stdout.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=["cfg2", "cfg3", "cfg5", "poisson", "unstructured"],
                    help="cfg2 = the main line (Quad4 4096^2); cfg3 = Quad9 2048^2 stiffness + mass; cfg5 = Tri3 shuffled; "
                         "poisson = cfg2's scalar one-dof variant (all three on 1 GPU)")
    ap.add_argument("--operator", default=None, choices=["reference", "poisson"], help="alias: --operator poisson = --config poisson")
    ap.add_argument("--n", "--nx", dest="n", type=int, default=4096, help="elements per side of each rank's block")
    ap.add_argument("--cg-iters", type=int, default=200)
    ap.add_argument("--cpu-sample", type=int, default=1024, help="side of the CPU-baseline sample block")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong_cfg4 leg")
    ap.add_argument("--strong-n", type=int, default=8192, help="side of the strong-scaling mesh (BASELINE configs[3]: 8192)")
    ap.add_argument("--full-n", type=int, default=512, help="side of the e2e_full mesh")
    ap.add_argument("--kernel", default="auto", choices=["auto", "ws", "phased", "rows", "tiled"], help="numeric assembly kernel")
    ap.add_argument("--halo", default="p2p", choices=["p2p", "nccl"], help="N > 1: per-iteration halo / scalar exchange")
    ap.add_argument("--ny-per-rank", type=int, default=0,
                    help="element rows per rank for N > 1 (default --n: weak scaling)")
    args = ap.parse_args()
    if args.kernel == "tiled":
        args.kernel = "auto"
    if args.operator == "poisson":
        args.config = "poisson"
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
