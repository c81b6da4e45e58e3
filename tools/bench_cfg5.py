#!/usr/bin/env python
"""BASELINE.json configs[4] (cfg5): 3-D ComplexF64 shifted-Laplacian Helmholtz, rediscretised on every level
(multilevelOperatorConstructor path, MGsetup.jl:28,105-106), V(2,2) cycle + FGMRES(5), row-partitioned into z-slabs
over N GPUs of one box (STRONG scaling: the same global grid for every N).

    python tools/bench_cfg5.py --cells 512 --levels 7                       # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
        tools/bench_cfg5.py --cells 512 --levels 7 --norms-ref profiles/r02_cfg5_n1_norms.json

Rank 0 prints one JSON line in the format of bench.py.  Parity: with --oracle (default for <= 256 cells) the per-cycle
residual norms of two solveMG cycles are compared with the CPU oracle on the same global hierarchy; --norms-out /
--norms-ref write / compare those norms across runs with different N (the right-hand side is seeded per node plane, so
every N solves the same system)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bench import ClockSampler, cycle_bytes_sizes, host_cores, log, measured_peaks  # noqa: E402


def plane_rhs(plane_lo, plane_hi, plane_len):
    """Right-hand side of node planes [plane_lo, plane_hi): seeded per plane, so any partition builds the same b."""
    out = np.empty((plane_hi - plane_lo) * plane_len, dtype=np.complex128)
    for k in range(plane_lo, plane_hi):
        rng = np.random.default_rng(7000 + k)
        v = rng.random(plane_len) + 1j * rng.random(plane_len)
        out[(k - plane_lo) * plane_len:(k - plane_lo + 1) * plane_len] = v
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=512)
    ap.add_argument("--levels", type=int, default=7)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--ppw", type=float, default=10.0, help="grid points per wavelength on the fine level")
    ap.add_argument("--oracle", type=int, default=-1, help="1: compare with the CPU oracle (global hierarchy on rank 0)")
    ap.add_argument("--norms-out", default="")
    ap.add_argument("--norms-ref", default="")
    ap.add_argument("--fgmres-restarts", type=int, default=20)
    args = ap.parse_args()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    import torch
    import multigrid_jl_b200 as mg
    import __graft_entry__ as g
    g.build_oracle()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    gloo = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        gloo = dist.new_group(backend="gloo")

    def gather(o):
        if world == 1:
            return [o]
        out = [None] * world
        dist.all_gather_object(out, o, group=gloo)
        return out
    cells, levels = args.cells, args.levels
    n = [cells] * 3
    dom = [0.0, 1.0, 0.0, 1.0, 0.0, 1.0]
    h = 1.0 / cells
    kappa2 = (2 * np.pi / (args.ppw * h)) ** 2
    op = mg.poisson_window_operator(dom, n, kappa2=kappa2, gamma=0.5)
    plane = (cells + 1) ** 2
    t0 = time.time()
    p = mg.getMGparam(np.complex128, np.int64, levels, 8, 20, 1e-6, "Jac", 0.8, 2, 2, 'V')
    p.nrhs = 1
    if world == 1:
        ctor = mg.getMultilevelOperatorConstructor(kappa2, lambda mesh, k2: op(np.asarray(mesh.n), 0, int(mesh.n[2]) + 1),
                                                   lambda mf, mc, pf, level: pf)
        mg.MGsetup(ctor, mg.getRegularMesh(dom, n), p, 1)
        log(f"[cfg5] host setup {time.time() - t0:.1f} s: rows {[a.shape[0] for a in p.As]}")
        t0 = time.time()
        dev = mg.DeviceHierarchy(p, device=local_rank)
        sizes = [(p.As[l].shape[0], p.As[l].nnz, p.As[l + 1].shape[0], p.Ps[l].nnz) for l in range(len(p.As) - 1)]
        b = plane_rhs(0, cells + 1, plane)
        lo_hi = (0, cells + 1)
    else:
        dh = mg.setup_slab_hierarchy(op, dom, n, p, rank, world, replicate_below=300000, gather=gather, rediscretise=True)
        log(f"[cfg5 rank {rank}] slab setup {time.time() - t0:.1f} s: {dh.nd} distributed + {len(dh.replicated.As)} replicated levels")
        t0 = time.time()
        ids = [mg.DeviceHierarchy.dist_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0, group=gloo)
        dev = mg.DeviceHierarchy.from_dist(dh, p, local_rank, ids[0])
        sizes = []
        for dl in dh.dist_levels:
            tot = np.sum(gather(np.array([dl.AT.nnz, dl.PT.nnz], dtype=np.int64)), axis=0)
            sizes.append((int(dl.n_global), int(tot[0]), int(dl.nc_global), int(tot[1])))
        rep = dh.replicated
        for j in range(len(rep.As) - 1):
            sizes.append((rep.As[j].shape[0], rep.As[j].nnz, rep.As[j + 1].shape[0], rep.Ps[j].nnz))
        lo_hi = mg.slab_planes(cells, world)[rank]
        b = plane_rhs(lo_hi[0], lo_hi[1], plane)
    N_total = (cells + 1) ** 3
    nb2 = sum(gather(float(np.vdot(b, b).real)))
    b /= np.sqrt(nb2)
    log(f"[cfg5 rank {rank}] upload {time.time() - t0:.1f} s, planes {lo_hi}")
    nbytes = cycle_bytes_sizes(sizes, sv=16)

    # ---- parity material: two solveMG cycles -------------------------------------------------------------------------
    x = np.zeros_like(b)
    _, it, res = dev.solveMG(b, x, 0.0, 2)
    res = np.array(res, copy=True)
    log(f"[cfg5 rank {rank}] residual norms of 2 cycles: {res.tolist()}")

    # ---- device-resident cycles ------------------------------------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        dev.cycle_device(True)
    dev.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = dev.launch_count()
    tw0 = time.time()
    dev.event_record(0)
    for _ in range(args.steps):
        dev.cycle_device(True)
    dev.event_record(1)
    ms = dev.event_elapsed_ms(0, 1)
    dev.synchronize()
    tw1 = time.time()
    launches = dev.launch_count() - l0
    clocks = sampler.stop(tw0, tw1)
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
        ms = float(t.item())
    ms_per_step = ms / args.steps
    dev.profile_enable(True)
    for _ in range(args.steps):
        dev.cycle_device(True)
    prof = dev.profile_report()
    dev.profile_enable(False)
    tot_ms = sum(r["total_ms"] for r in prof)
    kern = [{"kind": r["kind"], "level": r["level"], "launches": r["launches"], "avg_us": 1e3 * r["total_ms"] / r["launches"],
             "gbs": r["bytes"] / (r["total_ms"] * 1e-3) / 1e9 if r["total_ms"] > 0 else 0.0,
             "format_gbs": r["format_bytes"] / (r["total_ms"] * 1e-3) / 1e9 if r["total_ms"] > 0 else 0.0,
             "share": r["total_ms"] / tot_ms} for r in sorted(prof, key=lambda r: -r["total_ms"])]
    dom_k = max(prof, key=lambda r: r["total_ms"])
    hbm_peak, peak_src = measured_peaks()
    fmt_cycle = sum(r["format_bytes"] for r in prof) / args.steps
    achieved = dom_k["bytes"] / (dom_k["total_ms"] * 1e-3) / 1e9
    fmt_ach = dom_k["format_bytes"] / (dom_k["total_ms"] * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": f"{dom_k['kind']} level {dom_k['level']}", "achieved": achieved, "peak": hbm_peak,
                "unit": "GB/s", "frac": achieved / hbm_peak, "peak_source": peak_src, "traffic": None,
                "share_of_step": dom_k["total_ms"] / tot_ms, "format_achieved": fmt_ach, "format_frac": fmt_ach / hbm_peak,
                "cycle_algorithmic_gb": nbytes / 1e9, "cycle_achieved_gbs": nbytes / (ms_per_step * 1e-3) / 1e9,
                "cycle_frac": nbytes / (ms_per_step * 1e-3) / 1e9 / hbm_peak / world,
                "cycle_format_gb": fmt_cycle / 1e9 * (world if world > 1 else 1),
                "note": "per-GPU kernel figures; achieved/frac use the ALGORITHMIC CSR bytes of SURVEY.md 8(d) (16-byte values), "
                        "format_* the bytes the stencil-dictionary format really streams"}

    # ---- FGMRES(5) preconditioned by the cycle (solveGMRES_MG(As[1], param, b, x0, true, 5)) ----------------------------
    t0 = time.time()
    xk, itk, flag, resk = dev.solveFGMRES(b, np.zeros_like(b), 5, True, 1e-6, args.fgmres_restarts)
    t_fg = time.time() - t0
    fgmres = {"restarts": int(itk), "inner_steps": int(len(resk)), "flag": int(flag), "final_relres": float(resk[-1]) if len(resk) else None,
              "seconds_incl_host_copies": t_fg, "tol": 1e-6}
    log(f"[cfg5 rank {rank}] FGMRES(5): {len(resk)} inner steps, flag {flag}, {t_fg:.2f} s")

    # ---- end to end: one cycle per call through the host-buffer ABI ------------------------------------------------------
    import ctypes
    from multigrid_jl_b200.device import lib, _check
    nloc = dev.n
    hb = torch.empty(2 * nloc, dtype=torch.float64).pin_memory()
    hx = torch.empty(2 * nloc, dtype=torch.float64).pin_memory()
    hb.numpy().view(np.complex128)[:] = b

    def e2e_cycle():
        t0 = time.perf_counter()
        _check(lib().mgb200_precondition(dev.h, ctypes.c_void_p(hb.data_ptr()), ctypes.c_void_p(hx.data_ptr())))
        return time.perf_counter() - t0
    e2e_cycle()
    if world > 1:
        dist.barrier()
    te = sum(e2e_cycle() for _ in range(3)) / 3
    if world > 1:
        t = torch.tensor([te], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        te = float(t.item())
    e2e = {"value": N_total / te, "unit": "DOF/s", "h2d_bytes_per_step": 16 * N_total, "d2h_bytes_per_step": 16 * N_total,
           "ms_per_step": te * 1e3, "call": "mgb200_precondition, pinned host buffers, one V(2,2) cycle per call"}

    out = {"metric": "vcycle_dof_per_s", "value": N_total / (ms_per_step * 1e-3), "unit": "DOF/s", "n_gpus": world,
           "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": "c128", "data": "synthetic",
           "config": {"workload": f"cfg5: 3D ComplexF64 shifted-Laplacian Helmholtz {cells}^3 cells ({cells + 1}^3 nodes), "
                                  f"{args.ppw:g} points per wavelength, shift 0.5i, geometric MG rediscretised on {dev.levels} levels, "
                                  f"damped Jacobi 0.8, one V(2,2) cycle from x=0 per step",
                      "rows": N_total, "l2_policy": "inputs larger than L2",
                      "parallelism": f"row-partitioned z-slabs x{world}" if world > 1 else "single GPU"},
           "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "fgmres5": fgmres,
           "residual_norms_2_cycles": res.tolist(), "kernels": kern[:12]}

    # ---- parity ------------------------------------------------------------------------------------------------------------
    want_oracle = args.oracle == 1 or (args.oracle < 0 and cells <= 256)
    if rank == 0 and want_oracle:
        from oracle import cycle as oc
        if world == 1:
            pg, bg = p, b
        else:
            pg = mg.getMGparam(np.complex128, np.int64, levels, 8, 20, 1e-6, "Jac", 0.8, 2, 2, 'V')
            ctor = mg.getMultilevelOperatorConstructor(kappa2, lambda mesh, k2: op(np.asarray(mesh.n), 0, int(mesh.n[2]) + 1),
                                                       lambda mf, mc, pf, level: pf)
            mg.MGsetup(ctor, mg.getRegularMesh(dom, n), pg, 1)
            bg = plane_rhs(0, cells + 1, plane) / np.sqrt(nb2)
        o = oc.OracleMG(pg, numCores=host_cores())
        o.maxOuterIter, o.relativeTol = 2, 0.0
        t0 = time.perf_counter()
        _, _, res_ref = oc.solveMG(o, bg, np.zeros_like(bg))
        tc = (time.perf_counter() - t0) / 2
        rel = [abs(res[k] - res_ref[k]) / res_ref[k] for k in range(3)]
        out["parity"] = {"max_rel": float(max(rel)), "tol": 1e-10, "rel": [float(v) for v in rel],
                         "what": "per-cycle residual norms of solveMG (2 cycles) vs the CPU oracle on the global hierarchy"}
        out["cpu_baseline"] = {"value": N_total / tc, "unit": "DOF/s", "cores": host_cores(), "kind": "port",
                               "sample": "2 solveMG cycles (cycle + residual) of the same hierarchy on the host"}
        assert max(rel) <= 1e-10, out["parity"]
    if rank == 0 and args.norms_out:
        with open(args.norms_out, "w") as f:
            json.dump({"cells": cells, "levels": levels, "ppw": args.ppw, "n_gpus": world, "residual_norms": res.tolist(),
                       "fgmres_inner_steps": int(len(resk)), "fgmres_resvec": [float(v) for v in resk]}, f)
    if rank == 0 and args.norms_ref:
        with open(args.norms_ref) as f:
            ref = json.load(f)
        assert ref["cells"] == cells and ref["levels"] == levels
        rr = np.array(ref["residual_norms"])
        rel = np.abs(res - rr) / rr
        out["parity_vs_run"] = {"reference_run": os.path.basename(args.norms_ref), "reference_n_gpus": ref["n_gpus"],
                                "max_rel": float(rel.max()), "tol": 1e-10,
                                "fgmres_inner_steps": [int(len(resk)), int(ref["fgmres_inner_steps"])]}
        assert rel.max() <= 1e-10 and len(resk) == ref["fgmres_inner_steps"], out["parity_vs_run"]
    if rank == 0:
        os.write(json_fd, (json.dumps(out) + "\n").encode())
    dev.destroy()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
