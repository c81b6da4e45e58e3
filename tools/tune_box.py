#!/usr/bin/env python
"""Cycle time and per-kernel times of the bench workload (cfg2) for the variants of the box-stencil kernel
(csrc/box.cuh; options "box", "box_variant", "box_min_rows").  Results must not change by a bit."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import multigrid_jl_b200 as mg  # noqa: E402
from bench import build_problem  # noqa: E402
from tune_tma import timeit  # noqa: E402


def main():
    cells = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    levels = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    A, M, p, b = build_problem(cells, levels)
    dev = mg.DeviceHierarchy(p, device=0)
    x = np.zeros_like(b)
    dev.set_option("box", 0)
    _, _, res0 = dev.solveMG(b, x, 0.0, 2)
    ms, kern = timeit(dev)
    print(json.dumps({"box": 0, "cycle_ms": round(ms, 4), "kernels_us": kern}), flush=True)
    for variant in range(8):
        for minrows in (100000,):
            dev.set_option("box", 1)
            dev.set_option("box_variant", variant)
            dev.set_option("box_min_rows", minrows)
            _, _, res = dev.solveMG(b, x, 0.0, 2)
            same = bool(np.array_equal(res, res0))
            ms, kern = timeit(dev)
            print(json.dumps({"box": 1, "variant": variant, "min_rows": minrows, "bit_identical": same,
                              "cycle_ms": round(ms, 4), "kernels_us": kern}), flush=True)
    dev.destroy()


if __name__ == "__main__":
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    main()
