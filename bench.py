#!/usr/bin/env python3
"""Benchmark of the FEM hot path on B200 (contract: one JSON line on stdout from rank 0).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--n NX]

Workload (BASELINE.json configs[1]): Quad4 structured mesh NX x NX (default 4096 -> 16.8 M elements),
2-dof reference operator, fp64.  For N > 1 every rank owns one NX x NX element block of an
NX x (N*NX) mesh (weak scaling): ghost-element assembly of its own rows, halo-exchange PCG.

A "step" = one pass of numeric assembly over the rank's elements (element matrices evaluated on the
fly + deterministic CSR reduction) + boundary-condition application.  `value` = elements assembled
per second over all ranks with inputs resident in HBM.  The symbolic phase is one-time per mesh and
reported separately.  The CG leg (SpMV GB/s, PCG iterations/s) is timed in the same run and reported
under "cg"; `roofline` describes the assembly kernel and `cg.roofline` the SpMV kernel.

`e2e` is the same metric through the public API (FemProblem ... assemble()) on a fresh problem with
HOST (pinned) connectivity/coordinates each step: host->device copies, tag conversion, symbolic phase,
numeric assembly, BCs, and the device->host read of the right-hand side + a checksum of K.
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
    process pool (the port itself is single-threaded NumPy/SciPy, like the reference).  Returns
    (elements/s over all processes, wall seconds)."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    with ctx.Pool(procs) as pool:
        pool.map(_cpu_worker, [32] * procs)                 # start the workers, import numpy/scipy
        t0 = time.perf_counter()
        pool.map(_cpu_worker, [sample_n] * procs)
        dt = time.perf_counter() - t0
    return procs * sample_n * sample_n / dt, dt


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
                   "sample": "%d blocks of %dx%d elements per step, one per host core" % (procs, sample_n, sample_n)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "host_cores": os.cpu_count(), "kind": "port",
                         "sample": "oracle/fe_oracle.py assemble_csr + apply_bc_reference on %d independent %dx%d Quad4 "
                                   "blocks (%d elements) per step, one process per host core"
                                   % (procs, sample_n, sample_n, procs * sample_n * sample_n)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
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

    # ---- inputs resident in HBM -------------------------------------------------------------------
    if world == 1:
        conn_h, coords_h, bc = mesh.structured_quad4(n, dtype=np.int32)
        part = None
        from feprogram_b200.elements import _ELEM_CODES  # noqa: F401
        conn = torch.from_numpy(conn_h - 1).to(dev)
        coords = torch.from_numpy(coords_h).to(dev)
        nnode = coords_h.shape[0]
        t0 = time.perf_counter()
        pat = ops.symbolic(conn, nnode, 4, 2)
        plan = ops.tile_plan(conn, coords, pat) if args.kernel == "tiled" else None
        torch.cuda.synchronize()
        sym_times = []                          # rebuilds: exclude CUDA context / module load / first cudaMallocs
        for _ in range(3):
            del pat, plan
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            pat = ops.symbolic(conn, nnode, 4, 2)
            plan = ops.tile_plan(conn, coords, pat) if args.kernel == "tiled" else None
            torch.cuda.synchronize()
            sym_times.append(time.perf_counter() - t0)
        symbolic_s = sorted(sym_times)[1]
        nelem_local = conn.shape[0]
        owned_rows = pat.nrows
        d_dofs, n_dofs = [np.unique(np.concatenate([2 * (np.array(sorted(s), dtype=np.int64) - 1),
                                                     2 * (np.array(sorted(s), dtype=np.int64) - 1) + 1]))
                          for s in (bc[0], bc[3])]
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
    else:
        from feprogram_b200 import dist as fdist
        part = fdist.build_structured_block(n, rank, world, dev, args.ny_per_rank or n)
        symbolic_s = part.symbolic_s
        nelem_local = part.nelem_owned
        step = part.assemble_step
        launches_per_step = part.launches_per_step
        asm_bytes = part.asm_bytes
        spmv_bytes = part.spmv_bytes

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
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(K, 5)
        s0.record()
        for _ in range(reps):
            torch.ops.feprogram.spmv(pat.row_ptr, pat.col_idx, vals, x, y, None)
        s1.record()
        barrier()
        ms_spmv = s0.elapsed_time(s1) / reps
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
    if world == 1:
        conn1 = torch.from_numpy(conn_h).pin_memory()          # 1-based tags, host pinned
        xy_p = torch.from_numpy(coords_h).pin_memory()
        conn1_np, xy_np = conn1.numpy(), xy_p.numpy()
        out_rhs = torch.empty(pat.nrows, dtype=torch.float64).pin_memory()

        def e2e_step():
            fem = FemProblem(conn1_np, xy_np, device=dev)
            fem.setMatrix()
            fem.dofDir, fem.dofNeu = d_dofs, n_dofs            # what setBoundaryConditions computes (host sets)
            fem.assemble()
            st = fem._dev
            out_rhs.copy_(st["rhs"], non_blocking=True)
            chk = st["vals"].sum().item()                        # forces completion; 8-byte D2H
            return chk
        reps = max(1, min(K, 3))
        e2e_step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            chk = e2e_step()
        torch.cuda.synchronize()
        e2e_s = (time.perf_counter() - t0) / reps
        e2e = {"value": nelem_total / e2e_s, "unit": UNIT,
               "h2d_bytes_per_step": int(conn_h.nbytes + coords_h.nbytes + d_dofs.nbytes + n_dofs.nbytes),
               "d2h_bytes_per_step": int(pat.nrows * 8 + 8 + 8),
               "ms_per_step": 1e3 * e2e_s, "includes": "H2D, tag conversion, symbolic phase, numeric assembly, BCs, D2H rhs + checksum",
               "checksum": chk}
    elif part is not None:
        e2e = part.e2e(K)

    # ---- CPU baseline (rank 0, N == 1 only) -----------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        procs = os.cpu_count() or 1
        sn = min(args.cpu_sample, 512)
        v, dt = cpu_assembly_pool(sn, procs)
        cpu = {"value": v, "unit": UNIT, "cores": procs, "host_cores": os.cpu_count(), "kind": "port",
               "sample": "oracle/fe_oracle.py assemble_csr + apply_bc_reference on %d independent %dx%d Quad4 blocks "
                         "(%d elements, %.1f s wall), one process per host core" % (procs, sn, sn, procs * sn * sn, dt)}

    kernel_name = ("k_assemble_tiled" if (part is None and plan is not None) or (part is not None and part.tiled)
                   else "k_assemble_rows")
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "r01_dram_traffic.json")
    if os.path.exists(tpath) and world == 1 and n == 4096:      # ncu capture of exactly this workload
        with open(tpath) as f:
            traffic = json.load(f)
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
                       "symbolic_phase_s": symbolic_s, "partition": "dp%d element row blocks, owned-row assembly with ghost elements" % world},
            "roofline": {"kernel": kernel_name, "bound": "hbm", "achieved": asm_gbs, "peak": hbm_peak,
                         "unit": "GB/s", "frac": asm_gbs / hbm_peak,
                         "traffic": traffic.get("k_assemble_tiled") if kernel_name == "k_assemble_tiled" else None,
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": asm_bytes_total / world, "ms_per_launch": ms_kernel},
            "cg": {"spmv_ms": ms_spmv, "spmv_gbs": spmv_gbs, "spmv_frac": spmv_gbs / hbm_peak,
                   "pcg_iters_timed": pcg_done, "halo": (part.halo if part is not None else None), "pcg_iters_per_s": pcg_done / (ms_pcg * 1e-3) if ms_pcg > 0 else None,
                   "roofline": {"kernel": "k_spmv", "bound": "hbm", "achieved": spmv_gbs, "peak": hbm_peak, "unit": "GB/s",
                                "frac": spmv_gbs / hbm_peak, "traffic": traffic.get("k_spmv"),
                                "algorithmic_bytes_per_launch": spmv_bytes_total / world}},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches_per_step * K, "clocks": clk.summary(),
        }
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", "--nx", dest="n", type=int, default=4096, help="elements per side of each rank's block")
    ap.add_argument("--cg-iters", type=int, default=200)
    ap.add_argument("--cpu-sample", type=int, default=1024, help="side of the CPU-baseline sample block")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--kernel", default="tiled", choices=["tiled", "rows"], help="numeric assembly kernel")
    ap.add_argument("--ny-per-rank", type=int, default=0,
                    help="element rows per rank for N > 1 (default --n: weak scaling; 8192/N with --n 8192 = cfg4 strong scaling)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
