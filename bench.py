#!/usr/bin/env python
"""Benchmark of the multigrid hot path (BASELINE.json metric: V-cycle time & DOF/s, 256^3
Poisson; SpMV GB/s vs HBM peak).

    python bench.py --gpus 1 --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --steps K --warmup W    # restated reference CPU path

One step = one V(2,2) cycle of the geometric-MG hierarchy (Galerkin, 6 levels, damped Jacobi
0.8) on the 257^3-node Poisson problem of SURVEY.md section 8(d) cfg2, started from x = 0 as
the preconditioner closure does (SolveFuncs.jl:59).  `value` is DOF/s with everything
resident in HBM; `e2e` is the same metric through the host-buffer C-ABI call mgb200_solveMG
(b and x0 copied host->device, x copied back, per-cycle residual norms computed as the
reference's solveMG does) with pinned host buffers.
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


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ---------------------------------------------------------------------------------------------
# workload
# ---------------------------------------------------------------------------------------------
def build_problem(cells, levels, nrhs=1, seed=0):
    import multigrid_jl_b200 as mg
    t0 = time.time()
    M = mg.getRegularMesh([0, 1, 0, 1, 0, 1], [cells] * 3)
    A = mg.poisson_shifted(M, 1e-4)
    p = mg.getMGparam(np.float64, np.int64, levels, 8, 20, 1e-8, "Jac", 0.8, 2, 2, 'V')
    mg.MGsetup(A, M, p, nrhs)
    rng = np.random.default_rng(seed)
    shape = (A.shape[0],) if nrhs == 1 else (A.shape[0], nrhs)
    b = np.asfortranarray(A @ rng.random(shape))
    b /= np.linalg.norm(b)
    log(f"[bench] host setup {cells}^3 cells, {p.levels} levels: {time.time() - t0:.1f} s, "
        f"rows {[a.shape[0] for a in p.As]}, nnz {[a.nnz for a in p.As]}")
    return A, M, p, b


def weak_scaling_grid(cells, world, layout):
    """Cells per dimension of the N-GPU weak-scaling workload (cells^3 cells per GPU).  "cube": the grid doubles one
    dimension at a time, z first - 256x256x512 (N=2), 256x512x512 (N=4), 512^3 (N=8, the north-star size) - and
    is always cut into z-slabs; "stack": cells x cells x cells*N."""
    if layout == "cube" and world in (1, 2, 4, 8):
        f = {1: (1, 1, 1), 2: (1, 1, 2), 4: (1, 2, 2), 8: (2, 2, 2)}[world]
        return [cells * f[0], cells * f[1], cells * f[2]]
    return [cells, cells, cells * world]


def build_distributed(cells, levels, rank, world, local_rank, gloo_group, layout="cube", grid=None):
    """Weak-scaling workload for N > 1 (weak_scaling_grid): z-slab row partition following
    getOriginalBoundingBoxCells with NumCells = [1,1,N], Galerkin hierarchy built slab-locally on the host,
    coarse levels replicated."""
    import torch.distributed as dist
    import multigrid_jl_b200 as mg
    t0 = time.time()
    n = list(grid) if grid else weak_scaling_grid(cells, world, layout)
    dom = [0, n[0] / cells, 0, n[1] / cells, 0, n[2] / cells]
    if min(n) > cells and not grid:
        levels += 1          # every dimension doubled: one more level reaches the same coarsest grid
    p = mg.getMGparam(np.float64, np.int64, levels, 8, 20, 1e-8, "Jac", 0.8, 2, 2, 'V')
    p.nrhs = 1

    def gather(o):
        out = [None] * world
        dist.all_gather_object(out, o, group=gloo_group)
        return out
    dh = mg.setup_slab_hierarchy(mg.poisson_window_operator(dom, n, 1e-4), dom, n, p, rank, world,
                                 replicate_below=300000, gather=gather)
    ids = [mg.DeviceHierarchy.dist_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0, group=gloo_group)
    log(f"[bench rank {rank}] slab setup {time.time() - t0:.1f} s: {dh.nd} distributed + {len(dh.replicated.As)} "
        f"replicated levels, owned rows {[dl.AT.shape[1] for dl in dh.dist_levels]}")
    t0 = time.time()
    dev = mg.DeviceHierarchy.from_dist(dh, p, local_rank, ids[0])
    # global sizes for the byte model
    sizes = []
    for dl in dh.dist_levels:
        loc = np.array([dl.AT.nnz, dl.PT.nnz], dtype=np.int64)
        allv = gather(loc)
        tot = np.sum(allv, axis=0)
        sizes.append((int(dl.n_global), int(tot[0]), int(dl.nc_global), int(tot[1])))
    rep = dh.replicated
    for j in range(len(rep.As) - 1):
        sizes.append((rep.As[j].shape[0], rep.As[j].nnz, rep.As[j + 1].shape[0], rep.Ps[j].nnz))
    rng = np.random.default_rng(rank)
    b = rng.random(dev.n)
    nb2 = gather(float(np.dot(b, b)))
    b /= np.sqrt(sum(nb2))
    log(f"[bench rank {rank}] upload {time.time() - t0:.1f} s")
    return dev, p, b, sizes, int(dh.dist_levels[0].n_global), n


def cycle_bytes_sizes(sizes, nrhs=1, pre=2, post=2, sv=8):
    total, m = 0.0, nrhs
    for (n, nnz, nc, nnzP) in sizes:
        sweep = nnz * (sv + 4) + 4 * (n + 1) + (3 * m + 1) * n * sv
        resid = nnz * (sv + 4) + 4 * (n + 1) + 3 * n * sv * m
        restrict = nnzP * 12 + 4 * (nc + 1) + (n + nc) * sv * m
        prolong = nnzP * 12 + 4 * (n + 1) + (nc + 2 * n) * sv * m
        first = (2 * m + 1) * n * sv
        total += first + (pre - 1) * sweep + resid + restrict + prolong + post * sweep
    return total


def cycle_bytes(p, nrhs=1, pre=2, post=2, sv=8):
    """Algorithmic bytes of one V(pre,post) cycle from x = 0."""
    total = 0.0
    per_level = []
    m = nrhs
    for l in range(len(p.As) - 1):
        n, nnz = p.As[l].shape[0], p.As[l].nnz
        nc, nnzP = p.As[l + 1].shape[0], p.Ps[l].nnz
        sweep = nnz * (sv + 4) + 4 * (n + 1) + (3 * m + 1) * n * sv
        resid = nnz * (sv + 4) + 4 * (n + 1) + 3 * n * sv * m
        restrict = nnzP * 12 + 4 * (nc + 1) + (n + nc) * sv * m
        prolong = nnzP * 12 + 4 * (n + 1) + (nc + 2 * n) * sv * m
        first = (2 * m + 1) * n * sv
        lvl = first + (pre - 1) * sweep + resid + restrict + prolong + post * sweep
        per_level.append(lvl)
        total += lvl
    return total, per_level


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons DURING the timed region (NVML, every few ms)."""

    def __init__(self, gpu_index=0):
        self.rows = []
        self.gpu_index = gpu_index
        self.stop_flag = False
        self.thread = None
        self.err = None

    def _run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.gpu_index)
            self.smax = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                nv.nvmlDeviceGetCurrentClocksThrottleReasons
            while not self.stop_flag:
                self.rows.append((time.time(), nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), int(get_reasons(h)),
                                  nv.nvmlDeviceGetPowerUsage(h) / 1000.0))
                time.sleep(0.004)
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def start(self):
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def stop(self, t0=None, t1=None):
        self.stop_flag = True
        if self.thread is not None:
            self.thread.join(timeout=2.0)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + str(self.err)]}
        rows = [r for r in self.rows if (t0 is None or r[0] >= t0) and (t1 is None or r[0] <= t1)]
        if not rows:
            rows = self.rows
        bits = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
        reasons = set()
        for r in rows:
            for bit, nm in bits.items():
                if r[2] & bit:
                    reasons.add(nm)
        return {"sm_mhz": float(np.median([r[1] for r in rows])), "sm_max_mhz": float(self.smax),
                "reasons": sorted(reasons), "samples": len(rows), "power_w_max": max(r[3] for r in rows)}


def ncu_traffic(cells, dom, fmt):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture of the
    device format in use (fmt: "pattern" = stencil dictionary, "csr" = CSR stream)."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            t = json.load(f)
        return t[f"cfg2_{cells}"][f"{dom['kind']}_level{dom['level']}_{fmt}"]["traffic_bytes"]
    except Exception:
        return None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------
# CPU baseline (restated reference path = oracle)
# ---------------------------------------------------------------------------------------------
def host_cores():
    """Host threads the CPU arm uses: every core this process may run on.  Stated explicitly because
    torch.distributed.run exports OMP_NUM_THREADS=1 (the oracle's kernels take the thread count as an argument, like
    the numCores of the reference's SpMatMul, SpMatMul.jl:4)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_cycles(p, b, ncycles, nthreads=None, keep=None):
    """Time `ncycles` preconditioner-style V-cycles (z = 0; recursiveCycle) of the unfused,
    reference-order CPU restatement on all host cores.  keep: dict that receives the oracle and z of the last cycle."""
    from oracle import cycle as oc
    nthreads = host_cores() if nthreads is None else int(nthreads)
    o = oc.OracleMG(p, numCores=nthreads)
    MMG = oc.getMultigridPreconditioner(o, b)
    z = MMG(b)  # warm-up (page faults, thread pool)
    times = []
    for _ in range(ncycles):
        t0 = time.perf_counter()
        z = MMG(b)
        times.append(time.perf_counter() - t0)
    if keep is not None:
        keep["oracle"], keep["z"] = o, np.array(z, copy=True)
    return times, nthreads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cells, levels = args.cells, args.levels
    A, M, p, b = build_problem(cells, levels)
    N = A.shape[0]
    # the oracle's warm-up cycle is always run; further warm-up cycles as asked (they only add wall time)
    times, cores = cpu_cycles(p, b, max(args.steps, 1) + max(args.warmup - 1, 0))
    times = times[max(args.warmup - 1, 0):]
    tmax = float(np.mean(times))
    val = N / tmax
    nbytes, _ = cycle_bytes(p)
    workload = (f"cfg2: 3D Poisson {cells}^3 cells ({cells + 1}^3 nodes), geometric MG Galerkin "
                f"{p.levels} levels, damped Jacobi 0.8, one V(2,2) cycle from x=0 per step")
    if args.gpus > 1:
        g = weak_scaling_grid(cells, args.gpus, args.layout)
        workload = (f"BOUNDED SAMPLE of the {args.gpus}-GPU weak-scaled workload ({g[0]}x{g[1]}x{g[2]} cells): the CPU arm "
                    f"times the per-GPU share, " + workload + "; DOF/s of this memory-bound CPU path does not grow with "
                    "the grid, so the value stands for the whole grid on the same host")
    out = {
        "impl": "reference", "metric": "vcycle_dof_per_s", "value": val, "unit": "DOF/s", "n_gpus": args.gpus,
        "steps": len(times), "warmup": max(args.warmup, 1), "ms_per_step": tmax * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload,
                   "rows": N, "parallelism": f"host CPU, {cores} OpenMP threads (set explicitly)"},
        "cpu_baseline": {"value": val, "unit": "DOF/s", "cores": cores, "kind": "port",
                         "sample": f"{len(times)} full V(2,2) cycles of the {cells + 1}^3 hierarchy (unfused reference "
                                   f"order, OpenMP row-parallel SpMV, Int64 indices)"
                                   + ("" if args.gpus == 1 else f"; bounded sample of the {args.gpus}-GPU weak-scaled "
                                      f"workload: DOF/s of the memory-bound CPU path does not depend on the grid size"),
                         "effective_gbs": nbytes / tmax / 1e9},
        "e2e": {"value": val, "unit": "DOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def run_ours(args):
    # Libraries (NCCL's version banner, ...) may write to stdout: keep file descriptor 1 for the ONE JSON line and
    # send everything else to stderr.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    import torch
    import multigrid_jl_b200 as mg

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the solve phase has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    cells, levels = args.cells, args.levels
    if world > 1:
        gloo = dist.new_group(backend="gloo")
        dev, p, b, sizes, N, grid = build_distributed(cells, levels, rank, world, local_rank, gloo, args.layout,
                                                           [int(v) for v in args.grid.split(",")] if args.grid else None)
        nbytes_cycle = cycle_bytes_sizes(sizes)
        workload = (f"cfg2 weak-scaled to {world} GPUs: 3D Poisson {grid[0]}x{grid[1]}x{grid[2]} cells "
                    f"({grid[0] * grid[1] * grid[2] / world / cells ** 3:.3g} x {cells}^3 per GPU), "
                    f"z-slab row partition ({grid[2] // world} planes of {grid[0] + 1}x{grid[1] + 1} nodes per GPU), geometric "
                    f"MG Galerkin {dev.levels} levels, damped Jacobi 0.8, one V(2,2) cycle from x=0 per step; halo "
                    f"exchange + coarse gather inside the cycle")
        N_total = N
    else:
        A, M, p, b = build_problem(cells, levels, seed=rank)
        N = A.shape[0]
        t0 = time.time()
        dev = mg.DeviceHierarchy(p, device=local_rank)
        p.device = dev
        nbytes_cycle = cycle_bytes(p)[0]
        workload = (f"cfg2: 3D Poisson {cells}^3 cells ({cells + 1}^3 nodes), geometric MG Galerkin "
                    f"{p.levels} levels, damped Jacobi 0.8, one V(2,2) cycle from x=0 per step")
        N_total = N
        log(f"[bench] upload {time.time() - t0:.1f} s; kernel config level 1 A: {dev.kernel_config(1, 0)}, "
            f"P: {dev.kernel_config(1, 1)}, R: {dev.kernel_config(1, 2)}; level 2 A: {dev.kernel_config(2, 0)}")

    # GPU side of the parity record (compared with the oracle on the SAME full-size hierarchy further down, where the
    # cpu_baseline leg runs): two cycles of solveMG - per-cycle residual norms and the iterate - and the z of one
    # preconditioner cycle, all through the host-buffer C ABI
    x = np.zeros_like(b)
    xx, it, res = dev.solveMG(b, x, 0.0, 2)
    assert res[2] < res[1] < res[0], "cycle does not reduce the residual"
    gpu_par = {"res": np.array(res, copy=True), "xnorm": float(np.linalg.norm(xx)),
               "znorm": float(np.linalg.norm(dev.precondition(b)))}
    log(f"[bench rank {rank}] relres after 1,2 cycles: {res[1] / res[0]:.4e} {res[2] / res[0]:.4e}")

    # ---- device-resident V-cycles ------------------------------------------------------------
    db, dx = dev.device_buffers()  # b is already resident from the solve above
    for _ in range(max(args.warmup, 3)):
        dev.cycle_device(True)
    dev.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = dev.launch_count()
    from multigrid_jl_b200.device import lib as _lib
    _lib().mgb200_profiler_start()   # `ncu --profile-from-start off` captures exactly the timed region
    tw0 = time.time()
    dev.event_record(0)
    for _ in range(args.steps):
        dev.cycle_device(True)
    dev.event_record(1)
    ms = dev.event_elapsed_ms(0, 1)
    dev.synchronize()
    tw1 = time.time()
    _lib().mgb200_profiler_stop()
    launches = dev.launch_count() - launches0
    clocks = sampler.stop(tw0, tw1)
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = N_total / (ms_per_step * 1e-3)

    # ---- per-kernel timing of the same steps (CUDA events on the launching stream) --------------
    dev.profile_enable(True)
    for _ in range(args.steps):
        dev.cycle_device(True)
    prof = dev.profile_report()
    dev.profile_enable(False)
    hbm_peak, peak_src = measured_peaks()
    dom = max(prof, key=lambda r: r["total_ms"])
    tot_ms = sum(r["total_ms"] for r in prof)
    kern = []
    for r in sorted(prof, key=lambda r: -r["total_ms"]):
        gbs = r["bytes"] / (r["total_ms"] * 1e-3) / 1e9 if r["total_ms"] > 0 else 0.0
        fgbs = r["format_bytes"] / (r["total_ms"] * 1e-3) / 1e9 if r["total_ms"] > 0 else 0.0
        kern.append({"kind": r["kind"], "level": r["level"], "launches": r["launches"],
                     "avg_us": 1e3 * r["total_ms"] / r["launches"], "gbs": gbs, "format_gbs": fgbs,
                     "share": r["total_ms"] / tot_ms})
    log(f"[bench rank {rank}] per-kernel (events): " + json.dumps(kern[:8]))
    achieved = dom["bytes"] / (dom["total_ms"] * 1e-3) / 1e9
    fmt_achieved = dom["format_bytes"] / (dom["total_ms"] * 1e-3) / 1e9
    fmt_cycle = sum(r["format_bytes"] for r in prof) / args.steps
    pinfo = dev.pattern_info(1, 0)
    dinfo = dev.dist_info() if world > 1 else None
    nbytes = nbytes_cycle
    roofline = {"bound": "hbm", "kernel": f"{dom['kind']} level {dom['level']} (fused Jacobi sweep x' = x + d.*(b - A x))"
                if dom["kind"] == "sweep" else f"{dom['kind']} level {dom['level']}",
                "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "peak_source": peak_src,
                "traffic": ncu_traffic(cells, dom, "pattern" if pinfo["in_use"] else "csr"),
                "traffic_source": "constant from the committed ncu --set full capture of this kernel (profiles/ncu_traffic.json "
                                  "names the file); not measured by this run",
                "kernel_impl": "box_kernel (csrc/box.cuh: dense coefficient tables, TMA-staged plane windows, host-planned tile "
                               "records) on box-structured levels, else pat_tma_kernel / pat_kernel (csrc/pattern.cuh)",
                "share_of_step": dom["total_ms"] / tot_ms,
                "device_format": ("stencil dictionary (csrc/pattern.cuh, csrc/box.cuh): 16-bit pattern id per row, "
                                  f"{pinfo['patterns']} patterns / {pinfo['entries']} entries for A_1, d folded: "
                                  f"{pinfo['d_folded']}") if pinfo["in_use"] else "CSR (Int32 indices), TMA-staged",
                "format_bytes_per_launch": dom["format_bytes"] / dom["launches"],
                "format_achieved": fmt_achieved, "format_frac": fmt_achieved / hbm_peak,
                "note": "achieved/frac use the ALGORITHMIC CSR bytes of SURVEY.md 8(d) as the contract requires; "
                        "format_* use the bytes the device format really streams (frac > 1 means the kernel beats "
                        "the CSR-streaming roofline by not streaming the matrix)",
                "cycle_format_gb": fmt_cycle / 1e9,
                "cycle_format_gbs": fmt_cycle / (ms_per_step * 1e-3) / 1e9,
                "cycle_algorithmic_gb": nbytes / 1e9,
                "cycle_achieved_gbs": nbytes / (ms_per_step * 1e-3) / 1e9,
                "cycle_frac": nbytes / (ms_per_step * 1e-3) / 1e9 / hbm_peak / world}

    # ---- stand-alone fine-level SpMV (BASELINE metric "SpMV GB/s vs HBM peak") ---------------------
    sp_ms, sp_bytes = dev.bench_spmv(1, max(args.steps, 5))
    spmv = {"level": 1, "us": sp_ms * 1e3, "algorithmic_bytes": sp_bytes,
            "gbs": sp_bytes / (sp_ms * 1e-3) / 1e9, "frac_of_hbm_peak": sp_bytes / (sp_ms * 1e-3) / 1e9 / hbm_peak,
            "note": "y = A_1 x per GPU, CUDA events; algorithmic CSR bytes of SURVEY.md 8(d)"}

    # ---- end to end through the host-buffer C ABI ----------------------------------------------
    # (1) the per-step call: the closure of getMultigridPreconditioner (SolveFuncs.jl:43-63) = ONE V-cycle per call,
    #     r copied host -> device and z copied back inside the timed region.  This is `e2e.value`.
    # (2) a whole solveMG call of `cyc` cycles (b, x0 in; x out; per-cycle residual norms): copies amortised.
    cyc = args.e2e_cycles
    nloc = dev.n
    hb = torch.empty(nloc, dtype=torch.float64).pin_memory()
    hx = torch.empty(nloc, dtype=torch.float64).pin_memory()
    hb.numpy()[:] = b
    import ctypes
    from multigrid_jl_b200.device import lib, _check
    res = np.zeros(cyc + 1)
    itc = ctypes.c_int(0)

    def e2e_cycle():
        t0 = time.perf_counter()
        # synchronous: returns after the device -> host copy of z
        _check(lib().mgb200_precondition(dev.h, ctypes.c_void_p(hb.data_ptr()), ctypes.c_void_p(hx.data_ptr())))
        return time.perf_counter() - t0

    def e2e_solve():
        hx.zero_()       # the caller's x0 (the call overwrites x): prepared outside the timed call
        t0 = time.perf_counter()
        _check(lib().mgb200_solveMG(dev.h, ctypes.c_void_p(hb.data_ptr()), ctypes.c_void_p(hx.data_ptr()),
                                    ctypes.c_double(0.0), cyc, ctypes.byref(itc), res.ctypes.data_as(ctypes.c_void_p)))
        return time.perf_counter() - t0

    def timed(fn, reps):
        fn()
        if world > 1:
            dist.barrier()
        te = sum(fn() for _ in range(reps)) / reps
        if world > 1:
            t = torch.tensor([te], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            te = float(t.item())
        return te
    if world > 1:
        import torch.distributed as dist
    t_cycle = timed(e2e_cycle, max(3, min(args.steps, 10)))
    t_solve = timed(e2e_solve, max(2, min(args.steps, 5)))
    # what the host link gives a pinned copy of the same size (explains the gap between e2e and value)
    dtmp = torch.empty(nloc, dtype=torch.float64, device="cuda")
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dtmp.copy_(hb, non_blocking=True)
    torch.cuda.synchronize()
    ev0.record()
    dtmp.copy_(hb, non_blocking=True)
    ev1.record()
    torch.cuda.synchronize()
    h2d_gbs = nloc * 8 / (ev0.elapsed_time(ev1) * 1e-3) / 1e9
    del dtmp
    e2e = {"value": N_total / t_cycle, "unit": "DOF/s", "h2d_bytes_per_step": nloc * 8 * world,
           "d2h_bytes_per_step": nloc * 8 * world, "ms_per_step": t_cycle * 1e3,
           "call": "mgb200_precondition (the closure of getMultigridPreconditioner: z = one V(2,2) cycle from 0 on r), "
                   "pinned host buffers, ONE cycle per call, both copies inside the timed call",
           "host_link_h2d_gbs": h2d_gbs,
           "solve": {"value": N_total * cyc / t_solve, "unit": "DOF/s", "cycles_per_call": cyc, "ms_per_call": t_solve * 1e3,
                     "h2d_bytes_per_call": 2 * nloc * 8 * world, "d2h_bytes_per_call": (nloc * 8 + 8 * (cyc + 1)) * world,
                     "call": f"mgb200_solveMG (host buffers, pinned), {cyc} V(2,2) cycles per call incl. per-cycle "
                             f"residual norms: the copies are amortised over the cycles"}}

    out = {
        "metric": "vcycle_dof_per_s", "value": value, "unit": "DOF/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload, "rows": N_total,
                   "l2_policy": "inputs larger than L2 (fine-level vectors b, x, x', r: 4 x 136 MB per GPU touched every "
                                "cycle, plus the CSR arrays when the CSR-stream kernels run; L2 is 126 MB)",
                   "parallelism": f"row-partitioned z-slabs x{world}" if world > 1 else "single GPU",
                   "halo_exchange": (None if world == 1 else
                                     ("own kernels over NVLink peer memory (CUDA IPC, self-validating 8-byte words, one "
                                      "put+poll kernel per exchange); cycle incl. exchanges replayed from a CUDA graph"
                                      if dinfo["p2p"] else "ncclSend/ncclRecv"))},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "spmv": spmv,
        "kernel_options": {k: v for k, v in os.environ.items() if k.startswith("MGB200_")},
        "kernels": kern[:10],
    }
    if rank == 0 and world == 1 and not args.no_cpu:
        # bounded CPU sample: full cycles of the same hierarchy on all host cores
        keep = {}
        times, cores = cpu_cycles(p, b, args.cpu_cycles, keep=keep)
        tc = float(np.mean(times))
        out["cpu_baseline"] = {"value": N / tc, "unit": "DOF/s", "cores": cores, "kind": "port",
                               "sample": f"{len(times)} full V(2,2) cycles of the same {cells + 1}^3 hierarchy "
                                         f"(restated reference CPU path, unfused, OpenMP)",
                               "ms_per_cycle": tc * 1e3, "effective_gbs": nbytes / tc / 1e9}
        # parity of the TIMED configuration at full size against the oracle (north_star: per-cycle residual norms
        # within 1e-10 relative): solveMG for two cycles on the same hierarchy and b, plus the preconditioner's z
        from oracle import cycle as oc
        o = keep["oracle"]
        o.maxOuterIter, o.relativeTol = 2, 0.0
        x_ref, it_ref, res_ref = oc.solveMG(o, b, np.zeros_like(b))
        rel = [abs(gpu_par["res"][k] - res_ref[k]) / res_ref[k] for k in range(3)]
        rel.append(abs(gpu_par["xnorm"] - np.linalg.norm(x_ref)) / np.linalg.norm(x_ref))
        rel.append(abs(gpu_par["znorm"] - np.linalg.norm(keep["z"])) / np.linalg.norm(keep["z"]))
        out["parity"] = {"max_rel": float(max(rel)), "tol": 1e-10, "rows": int(N),
                         "what": "GPU (C ABI, host buffers) vs CPU oracle on the timed hierarchy: ||b||, the residual "
                                 "norms after cycles 1 and 2 of solveMG, ||x_2||, and ||z|| of one preconditioner cycle",
                         "rel": [float(v) for v in rel], "oracle": "restated reference path (parity unpinned: no golden "
                                                                   "vectors exist upstream, DESIGN.md section 3)"}
        if not max(rel) <= 1e-10:
            raise SystemExit(f"bench.py: parity against the oracle failed at full size: {out['parity']}")
    if rank == 0:
        os.write(json_fd, (json.dumps(out) + "\n").encode())
    os.close(json_fd)
    dev.destroy()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def run_config5(args, g):
    """cfg5 (BASELINE.json configs[4]) behind the same command line: our arm = tools/bench_cfg5.py (513^3 nodes, one
    global grid for every N); reference arm = the CPU oracle's V(2,2) cycle on a BOUNDED SAMPLE of the same family
    (129^3 nodes, same operator, wavelength per cell and smoother: DOF/s of the memory-bound CPU path does not depend on
    the grid size), all host cores, rank 0 only."""
    cells, levels = args.cfg5_cells, args.cfg5_levels
    if args.impl != "reference":
        if not os.path.exists(g.LIB):
            g.build_cuda()
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_cfg5
        sys.argv = ["bench_cfg5.py", "--cells", str(cells), "--levels", str(levels), "--steps", str(args.steps),
                    "--warmup", str(args.warmup)]
        # cross-N parity at the full size (the CPU oracle does not fit the time budget at 513^3): the per-cycle residual
        # norms and the FGMRES step count of the committed one-GPU run
        ref = os.path.join(ROOT, "profiles", "r02_cfg5_n1_norms.json")
        if cells == 512 and levels == 7 and int(os.environ.get("WORLD_SIZE", "1")) > 1 and os.path.exists(ref):
            sys.argv += ["--norms-ref", ref]
        bench_cfg5.main()
        return
    if int(os.environ.get("RANK", "0")) != 0:
        return
    g.build_oracle()
    import multigrid_jl_b200 as mg
    from oracle import cycle as oc
    sc, sl = 128, 5
    dom, n = [0.0, 1.0] * 3, [sc] * 3
    kappa2 = (2 * np.pi / (10.0 * (1.0 / sc))) ** 2
    p = mg.getMGparam(np.complex128, np.int64, sl, 8, 20, 1e-6, "Jac", 0.8, 2, 2, 'V')
    ctor = mg.getMultilevelOperatorConstructor(kappa2, lambda mesh, k2: mg.helmholtz_shifted(mesh, k2, 0.5),
                                               lambda mf, mc, pf, level: pf)
    mg.MGsetup(ctor, mg.getRegularMesh(dom, n), p, 1)
    N = p.As[0].shape[0]
    rng = np.random.default_rng(0)
    b = rng.random(N) + 1j * rng.random(N)
    b /= np.linalg.norm(b)
    cores = host_cores()
    o = oc.OracleMG(p, numCores=cores)
    MMG = oc.getMultigridPreconditioner(o, b)
    MMG(b)
    times = []
    for _ in range(max(args.steps, 1) + max(args.warmup - 1, 0)):
        t0 = time.perf_counter()
        MMG(b)
        times.append(time.perf_counter() - t0)
    times = times[max(args.warmup - 1, 0):]
    tm = float(np.mean(times))
    val = N / tm
    sample = (f"{len(times)} V(2,2) cycles of the cfg5 family at {sc + 1}^3 nodes (ComplexF64 shifted Laplacian, 10 points per "
              f"wavelength, rediscretised, {sl} levels): BOUNDED SAMPLE of the {cells + 1}^3 workload")
    out = {"impl": "reference", "metric": "vcycle_dof_per_s", "value": val, "unit": "DOF/s", "n_gpus": args.gpus,
           "steps": len(times), "warmup": max(args.warmup, 1), "ms_per_step": tm * 1e3, "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": "c128", "data": "synthetic",
           "config": {"workload": "cfg5: " + sample, "rows": N, "parallelism": f"host CPU, {cores} OpenMP threads (set explicitly)"},
           "cpu_baseline": {"value": val, "unit": "DOF/s", "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": val, "unit": "DOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells", type=int, default=256)
    ap.add_argument("--levels", type=int, default=6)
    ap.add_argument("--e2e-cycles", type=int, default=10)
    ap.add_argument("--cpu-cycles", type=int, default=5)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--grid", default="", help="N > 1: explicit cells per dimension n1,n2,n3 (development runs)")
    ap.add_argument("--layout", default="cube", choices=["cube", "stack"],
                    help="N > 1 weak-scaling grid: doubling dimensions up to 512^3 at N=8 (cube) or stacked in z")
    ap.add_argument("--config", type=int, default=2, choices=[2, 5],
                    help="BASELINE.json config: 2 (default, the metric's config: Float64 Poisson 256^3, weak-scaled for N > 1) or "
                         "5 (ComplexF64 Helmholtz 512^3 rediscretised, STRONG-scaled over N GPUs: tools/bench_cfg5.py)")
    ap.add_argument("--cfg5-cells", type=int, default=512)
    ap.add_argument("--cfg5-levels", type=int, default=7)
    args = ap.parse_args()
    import __graft_entry__ as g
    if args.config == 5:
        run_config5(args, g)
        return
    if args.impl == "reference":
        g.build_oracle()
        run_reference(args)
    else:
        if not os.path.exists(g.LIB):
            g.build_cuda()
        g.build_oracle()
        run_ours(args)


if __name__ == "__main__":
    main()
