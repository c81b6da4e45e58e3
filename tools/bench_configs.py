#!/usr/bin/env python
"""Measurements for the five BASELINE.json configs on one B200 (parity-test cases, not the bench
line): cycle time, DOF/s, achieved GB/s against the algorithmic byte model of SURVEY.md 8(d),
Krylov iterations and solve time, with the restated reference CPU path timed beside where it is
affordable.  Writes one JSON object per config to stdout / --out.

    python tools/bench_configs.py --configs 1,2,3,4,5 [--cfg3-cells 192] [--cfg5-cells 256]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import multigrid_jl_b200 as mg  # noqa: E402
from bench import measured_peaks  # noqa: E402


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def cycle_bytes(p, nrhs, pre, post, visits):
    sv = 16 if p.VAL == np.complex128 else 8
    m, total = nrhs, 0.0
    for l in range(len(p.As) - 1):
        n, nnz = p.As[l].shape[0], p.As[l].nnz
        nc, nnzP = p.As[l + 1].shape[0], p.Ps[l].nnz
        nnzR = p.Rs[l].nnz
        sweep = nnz * (sv + 4) + 4 * (n + 1) + (3 * m + 1) * n * sv
        resid = nnz * (sv + 4) + 4 * (n + 1) + 3 * n * sv * m
        restrict = nnzR * 12 + 4 * (nc + 1) + (n + nc) * sv * m
        prolong = nnzP * 12 + 4 * (n + 1) + (nc + 2 * n) * sv * m
        first = (2 * m + 1) * n * sv
        v = visits[l]
        # first visit starts from x = 0; revisits (W) start from x != 0: one more sweep-like pass
        lvl = first + (pre - 1) * sweep + resid + restrict + prolong + post * sweep
        re = pre * sweep + resid + restrict + prolong + post * sweep
        total += lvl * (v - v // 2 if v > 1 else 1) + re * (v // 2 if v > 1 else 0)
    return total


def time_cycles(dev, b, steps=20, warmup=5):
    x = np.zeros_like(b)
    dev.solveMG(b, x, 0.0, 1)           # puts b on the device
    for _ in range(warmup):
        dev.cycle_device(True)
    dev.synchronize()
    dev.event_record(0)
    for _ in range(steps):
        dev.cycle_device(True)
    dev.event_record(1)
    ms = dev.event_elapsed_ms(0, 1) / steps
    dev.profile_enable(True)
    for _ in range(steps):
        dev.cycle_device(True)
    prof = dev.profile_report()
    dev.profile_enable(False)
    tot = sum(r["total_ms"] for r in prof)
    kern = [{"kind": r["kind"], "level": r["level"], "launches": r["launches"] // steps,
             "avg_us": round(1e3 * r["total_ms"] / r["launches"], 2),
             "gbs": round(r["bytes"] / (r["total_ms"] * 1e-3) / 1e9, 1) if r["total_ms"] > 0 else 0,
             "share": round(r["total_ms"] / tot, 4)} for r in sorted(prof, key=lambda r: -r["total_ms"])[:8]]
    return ms, kern


def cpu_cycle_ms(p, b, n=2):
    from oracle import cycle as oc
    o = oc.OracleMG(p)
    MMG = oc.getMultigridPreconditioner(o, b)
    MMG(b)
    t0 = time.perf_counter()
    for _ in range(n):
        MMG(b)
    return (time.perf_counter() - t0) / n * 1e3, o.numCores


def run(name, A, AT, p, b, nrhs, pre, post, cyc, solver, inner=5, cpu=True, steps=20):
    hbm, src = measured_peaks()
    N = A.shape[0]
    t0 = time.time()
    dev = mg.uploadHierarchy(p)
    up = time.time() - t0
    L = len(p.As)
    visits = {'V': [1] * (L - 1), 'W': [2 ** l for l in range(L - 1)], 'F': [l + 1 for l in range(L - 1)]}[cyc]
    nbytes = cycle_bytes(p, nrhs, pre, post, visits)
    ms, kern = time_cycles(dev, b, steps=steps)
    rec = {"config": name, "rows": N, "nrhs": nrhs, "levels": L, "level_rows": [a.shape[0] for a in p.As],
           "level_nnz": [int(a.nnz) for a in p.As], "cycle": f"{cyc}({pre},{post})", "upload_s": round(up, 2),
           "cycle_ms": ms, "dof_per_s": N * nrhs / (ms * 1e-3), "cycle_algorithmic_gb": nbytes / 1e9,
           "cycle_gbs": nbytes / (ms * 1e-3) / 1e9, "cycle_frac_of_measured_hbm": nbytes / (ms * 1e-3) / 1e9 / hbm,
           "kernels": kern}
    x = np.zeros_like(b)
    t0 = time.perf_counter()
    if solver == "cg":
        x, _, it = mg.solveCG_MG(AT, p, b, x)
        res = p.last_resvec
        rec.update(solver="PCG" if nrhs == 1 else "blockCG", iters=it, flag=p.last_flag,
                   final_relres=float(np.max(res[-1])) if len(res) else None)
    else:
        x, _, it, res = mg.solveGMRES_MG(AT, p, b, x, True, inner)
        rec.update(solver=f"FGMRES({inner})", restarts=it, inner_steps=len(res), flag=p.last_flag,
                   final_relres=float(res[-1]) if len(res) else None)
    rec["solve_s"] = time.perf_counter() - t0
    rec["true_relres"] = float(np.linalg.norm(b - A @ x) / np.linalg.norm(b))
    if cpu:
        cms, cores = cpu_cycle_ms(p, b)
        rec["cpu_cycle_ms"] = cms
        rec["cpu_cores"] = cores
        rec["gpu_over_cpu"] = cms / ms
    mg.clear(p)
    return rec


def cfg1():
    M = mg.getRegularMesh([0, 1, 0, 1], [128, 128])
    A = mg.poisson_shifted(M, 1e-4)
    p = mg.getMGparam(np.float64, np.int64, 4, 8, 50, 1e-8, "Jac", 0.8, 2, 2, 'V')
    mg.MGsetup(A, M, p, 1)
    rng = np.random.default_rng(0)
    b = A @ rng.random(A.shape[0])
    b /= np.linalg.norm(b)
    return run("cfg1: 2D Poisson 128x128, GMG V(2,2), Jacobi, PCG", A, A, p, b, 1, 2, 2, 'V', "cg", steps=200)


def cfg2():
    M = mg.getRegularMesh([0, 1, 0, 1, 0, 1], [256] * 3)
    A = mg.poisson_shifted(M, 1e-4)
    p = mg.getMGparam(np.float64, np.int64, 6, 8, 50, 1e-8, "Jac", 0.8, 2, 2, 'V')
    mg.MGsetup(A, M, p, 1)
    rng = np.random.default_rng(0)
    b = A @ rng.random(A.shape[0])
    b /= np.linalg.norm(b)
    return run("cfg2: 3D Poisson 256^3, GMG V(2,2), Jacobi, PCG", A, A, p, b, 1, 2, 2, 'V', "cg", cpu=False)


def cfg3(cells, levels):
    rng = np.random.default_rng(0)
    M = mg.getRegularMesh([0, 1, 0, 1, 0, 1], [cells] * 3)
    w = mg.edge_weights_from_cells(M, np.exp(rng.standard_normal(cells ** 3)))
    A0 = mg.nodal_stencil_matrix(M, w, 0.0)
    A = mg.nodal_stencil_matrix(M, w, 1e-6 * abs(A0).sum(axis=0).max())
    del A0
    p = mg.getMGparam(np.float64, np.int64, levels, 8, 50, 1e-8, "SPAI", 1.0, 2, 2, 'W', "Julia", 0.4)
    t0 = time.time()
    mg.SA_AMGsetup(A, p, True, 1)
    log(f"cfg3 SA setup {time.time() - t0:.1f}s sizes {[a.shape[0] for a in p.As]} nnz {[a.nnz for a in p.As]}")
    b = A @ rng.random(A.shape[0])
    b /= np.linalg.norm(b)
    return run(f"cfg3: SA-AMG 3D diffusion {cells}^3, SPAI, W(2,2), FGMRES(5)", A, A, p, b, 1, 2, 2, 'W', "gmres",
               cpu=cells <= 128)


def cfg4():
    M = mg.getRegularMesh([0, 1, 0, 1, 0, 1], [128] * 3)
    A = mg.poisson_shifted(M, 1e-4)
    p = mg.getMGparam(np.float64, np.int64, 5, 8, 50, 1e-8, "Jac", 0.8, 2, 2, 'V')
    mg.MGsetup(A, M, p, 32)
    rng = np.random.default_rng(0)
    b = np.asfortranarray(A @ rng.random((A.shape[0], 32)))
    b /= np.linalg.norm(b)
    return run("cfg4: block MG 3D Poisson 128^3 x 32 RHS, V(2,2), blockCG", A, A, p, b, 32, 2, 2, 'V', "cg", cpu=False)


def cfg5(cells, levels):
    M = mg.getRegularMesh([0, 1, 0, 1, 0, 1], [cells] * 3)
    h = 1.0 / cells
    kappa2 = (2 * np.pi / (10 * h)) ** 2
    ctor = mg.getMultilevelOperatorConstructor(kappa2, lambda mesh, k2: mg.helmholtz_shifted(mesh, k2, 0.5),
                                               lambda mf, mc, pf, level: pf)
    p = mg.getMGparam(np.complex128, np.int64, levels, 8, 20, 1e-6, "Jac", 0.8, 2, 2, 'V')
    t0 = time.time()
    mg.MGsetup(ctor, M, p, 1)
    log(f"cfg5 setup {time.time() - t0:.1f}s rows {[a.shape[0] for a in p.As]}")
    AT = p.As[0]
    A = AT.conj().T.tocsr()
    rng = np.random.default_rng(0)
    b = A @ (rng.random(A.shape[0]) + 1j * rng.random(A.shape[0]))
    b /= np.linalg.norm(b)
    return run(f"cfg5: ComplexF64 shifted-Laplacian Helmholtz {cells}^3 (10 pts/wavelength), rediscretised GMG V(2,2), "
               f"FGMRES(5), 1 GPU", A, AT, p, b, 1, 2, 2, 'V', "gmres", cpu=False, steps=10)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,2,3,4,5")
    ap.add_argument("--cfg3-cells", type=int, default=128)
    ap.add_argument("--cfg3-levels", type=int, default=4)
    ap.add_argument("--cfg5-cells", type=int, default=256)
    ap.add_argument("--cfg5-levels", type=int, default=6)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    out = []
    for c in args.configs.split(","):
        t0 = time.time()
        rec = {"1": cfg1, "2": cfg2, "3": lambda: cfg3(args.cfg3_cells, args.cfg3_levels), "4": cfg4,
               "5": lambda: cfg5(args.cfg5_cells, args.cfg5_levels)}[c]()
        rec["wall_s"] = round(time.time() - t0, 1)
        print(json.dumps(rec), flush=True)
        out.append(rec)
    if args.out:
        with open(args.out, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
