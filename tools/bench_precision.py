#!/usr/bin/env python
"""cfg2 (3D Poisson 256^3 cells, Galerkin, V(2,2), Jacobi 0.8) with a Float64 and a Float32 hierarchy on one B200:
cycle time of each, PCG with the Float64 cycle against PCG in double precision over the Float32 cycle (the
reference's mixed-precision mode, SolveFuncs.jl:52-60).  One JSON object on stdout.

    python tools/bench_precision.py [--cells 256] [--levels 6]
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
from bench import build_problem, cycle_bytes, log  # noqa: E402


def cycle_ms(dev, steps=20, warmup=5):
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
    kern = {f"{r['kind']}{r['level']}": round(1e3 * r["total_ms"] / r["launches"], 1)
            for r in sorted(prof, key=lambda r: -r["total_ms"])[:8]}
    return ms, kern


def timed(f):
    t0 = time.perf_counter()
    out = f()
    return out, 1e3 * (time.perf_counter() - t0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=256)
    ap.add_argument("--levels", type=int, default=6)
    args = ap.parse_args()
    A, M, p64, b = build_problem(args.cells, args.levels)
    n = A.shape[0]
    out = {"workload": f"cfg2: 3D Poisson {args.cells}^3 cells ({n} rows), Galerkin, {p64.levels} levels, V(2,2), Jacobi 0.8"}

    # ---- Float64 -------------------------------------------------------------------------------------------------
    dev = mg.uploadHierarchy(p64)
    x = np.zeros_like(b)
    dev.solveMG(b, x, 0.0, 1)
    ms64, k64 = cycle_ms(dev)
    x = np.zeros_like(b)
    mg.solveCG_MG(A, p64, b, x)                     # warm-up (graph capture, workspaces)
    x = np.zeros_like(b)
    (_, _, it64), t64 = timed(lambda: mg.solveCG_MG(A, p64, b, x))
    res64 = np.linalg.norm(b - A @ x)
    out["float64"] = {"cycle_ms": ms64, "gdof_per_s": n / ms64 / 1e6, "kernels_us": k64, "pcg_iter": int(it64),
                      "pcg_ms_host_buffers": t64, "pcg_final_relres": float(res64),
                      "algorithmic_gb_per_cycle": cycle_bytes(p64)[0] / 1e9}
    log(f"[precision] float64: {json.dumps(out['float64'])}")
    mg.clear(p64)

    # ---- Float32 hierarchy ---------------------------------------------------------------------------------------
    t0 = time.time()
    p32 = mg.getMGparam(np.float32, np.int64, args.levels, 8, 20, 1e-8, "Jac", 0.8, 2, 2, 'V')
    mg.MGsetup(A, M, p32, 1)
    log(f"[precision] float32 host setup {time.time() - t0:.1f} s")
    dev = mg.uploadHierarchy(p32)
    bs = b.astype(np.float32)
    xs = np.zeros_like(bs)
    _, _, res = dev.solveMG(bs, xs, 0.0, 4)
    ms32, k32 = cycle_ms(dev)
    x = np.zeros_like(b)
    mg.solveCG_MG(A, p32, b, x)
    x = np.zeros_like(b)
    (_, _, it32), t32 = timed(lambda: mg.solveCG_MG(A, p32, b, x))
    resm = np.linalg.norm(b - A @ x)
    out["float32"] = {"cycle_ms": ms32, "gdof_per_s": n / ms32 / 1e6, "kernels_us": k32,
                      "solveMG_relres_per_cycle": [float(r / res[0]) for r in res],
                      "mixed_pcg_iter": int(it32), "mixed_pcg_ms_host_buffers": t32, "mixed_pcg_final_relres": float(resm)}
    log(f"[precision] float32: {json.dumps(out['float32'])}")
    out["cycle_speedup_float32"] = ms64 / ms32
    print(json.dumps(out))


if __name__ == "__main__":
    main()
