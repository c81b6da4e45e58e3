#!/usr/bin/env python
"""Cycle time and per-kernel times (CUDA events) of the bench workload for option sets given on the command line:
    python tools/tune.py [--cells 256 --levels 6] box=0 box=1,box_variant=1 grid_transfers=0 ...
Every set is applied with mgb200_set_option on top of the defaults; results must not change by a bit."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import multigrid_jl_b200 as mg  # noqa: E402
from bench import build_problem  # noqa: E402


def timeit(dev, steps=20, top=40):
    for _ in range(5):
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
            for r in sorted(prof, key=lambda r: -r["total_ms"])[:top]}
    return ms, kern


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=256)
    ap.add_argument("--levels", type=int, default=6)
    ap.add_argument("--helmholtz", action="store_true", help="cfg5's family: ComplexF64 shifted Laplacian, rediscretised")
    ap.add_argument("--grid", default="", help="Float64 Poisson on n1,n2,n3 cells instead of a cube (long lines: 512,512,64)")
    ap.add_argument("--nrhs", type=int, default=1, help="block variant: nrhs right-hand sides (cfg4: --cells 128 --levels 5 --nrhs 32)")
    ap.add_argument("sets", nargs="*")
    args = ap.parse_args()
    if args.helmholtz:
        M = mg.getRegularMesh([0, 1, 0, 1, 0, 1], [args.cells] * 3)
        kappa2 = (2 * np.pi / (10 * (1.0 / args.cells))) ** 2
        ctor = mg.getMultilevelOperatorConstructor(kappa2, lambda mesh, k2: mg.helmholtz_shifted(mesh, k2, 0.5),
                                                   lambda mf, mc, pf, level: pf)
        p = mg.getMGparam(np.complex128, np.int64, args.levels, 8, 20, 1e-6, "Jac", 0.8, 2, 2, 'V')
        mg.MGsetup(ctor, M, p, 1)
        rng = np.random.default_rng(0)
        b = rng.random(p.As[0].shape[0]) + 1j * rng.random(p.As[0].shape[0])
        b /= np.linalg.norm(b)
    elif args.grid:
        n = [int(v) for v in args.grid.split(",")]
        M = mg.getRegularMesh([0, n[0] / max(n), 0, n[1] / max(n), 0, n[2] / max(n)], n)
        A = mg.poisson_shifted(M, 1e-4)
        p = mg.getMGparam(np.float64, np.int64, args.levels, 8, 20, 1e-8, "Jac", 0.8, 2, 2, 'V')
        mg.MGsetup(A, M, p, 1)
        b = A @ np.random.default_rng(0).random(A.shape[0])
        b /= np.linalg.norm(b)
    else:
        A, M, p, b = build_problem(args.cells, args.levels, nrhs=args.nrhs)
    dev = mg.DeviceHierarchy(p, device=0)
    x = np.zeros_like(b)
    _, _, res0 = dev.solveMG(b, x, 0.0, 2)
    defaults = {}
    for spec in [""] + list(args.sets):
        opts = dict(kv.split("=") for kv in spec.split(",") if kv)
        for k, v in defaults.items():
            dev.set_option(k, v)
        for k, v in opts.items():
            defaults.setdefault(k, {"box": 1, "box_variant": -1, "box_variant27": -1, "box_variant_c": -1, "box_min_rows": 100000,
                                    "grid_transfers": 1, "tma": 1, "graphs": 1, "fuse_first": 1}.get(k, 0))
            dev.set_option(k, int(v))
        _, _, res = dev.solveMG(b, x, 0.0, 2)
        ms, kern = timeit(dev)
        print(json.dumps({"options": opts or "defaults", "bit_identical": bool(np.array_equal(res, res0)),
                          "cycle_ms": round(ms, 4), "kernels_us": kern}), flush=True)
    dev.destroy()


if __name__ == "__main__":
    main()
