#!/usr/bin/env python
"""Cycle time of the bench workload (cfg2) and of a ComplexF64 Helmholtz twin for different settings of the
TMA-staged dictionary kernel (options "tma", "tma_min_rows").  Results must not change."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import multigrid_jl_b200 as mg  # noqa: E402
from bench import build_problem  # noqa: E402


def timeit(dev, steps=20):
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
            for r in sorted(prof, key=lambda r: -r["total_ms"])[:12]}
    return ms, kern


def sweep(name, p, b):
    dev = mg.DeviceHierarchy(p, device=0)
    x = np.zeros_like(b)
    _, _, res0 = dev.solveMG(b, x, 0.0, 2)
    for tma, minrows in [(0, 0), (1, 0), (1, 200000), (1, 1000000), (1, 3000000)]:
        dev.set_option("tma", tma)
        dev.set_option("tma_min_rows", minrows)
        _, _, res = dev.solveMG(b, x, 0.0, 2)
        assert np.array_equal(res, res0)
        ms, kern = timeit(dev)
        print(json.dumps({"problem": name, "tma": tma, "tma_min_rows": minrows, "cycle_ms": round(ms, 4), "kernels_us": kern}),
              flush=True)
    dev.destroy()


def main():
    A, M, p, b = build_problem(256, 6)
    sweep("cfg2 257^3 Float64", p, b)
    del A, p, b
    cells = 192
    M = mg.getRegularMesh([0, 1, 0, 1, 0, 1], [cells] * 3)
    kappa2 = (2 * np.pi / (10 * (1.0 / cells))) ** 2
    ctor = mg.getMultilevelOperatorConstructor(kappa2, lambda mesh, k2: mg.helmholtz_shifted(mesh, k2, 0.5),
                                               lambda mf, mc, pf, level: pf)
    p = mg.getMGparam(np.complex128, np.int64, 5, 8, 20, 1e-6, "Jac", 0.8, 2, 2, 'V')
    mg.MGsetup(ctor, M, p, 1)
    A = p.As[0].conj().T.tocsr()
    rng = np.random.default_rng(0)
    b = A @ (rng.random(A.shape[0]) + 1j * rng.random(A.shape[0]))
    b /= np.linalg.norm(b)
    sweep("Helmholtz 193^3 ComplexF64", p, b)


if __name__ == "__main__":
    main()
