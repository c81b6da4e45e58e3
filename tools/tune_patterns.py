#!/usr/bin/env python
"""Sweep the launch parameters of the stencil-dictionary kernel (csrc/pattern.cuh) on the bench
workload (cfg2) and print cycle time + per-kernel timings for each setting.

    python tools/tune_patterns.py [--cells 256] [--rpt 1,2,4,8]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import multigrid_jl_b200 as mg  # noqa: E402
from bench import build_problem  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=256)
    ap.add_argument("--levels", type=int, default=6)
    ap.add_argument("--rpt", default="1,2,4,8")
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    A, M, p, b = build_problem(args.cells, args.levels)
    dev = mg.DeviceHierarchy(p, device=0)
    x = np.zeros_like(b)
    _, _, res0 = dev.solveMG(b, x, 0.0, 2)
    for rpt in [int(v) for v in args.rpt.split(",")]:
        dev.set_option("pattern_rows_per_thread", rpt)
        _, _, res = dev.solveMG(b, x, 0.0, 2)
        assert np.array_equal(res, res0), "rows-per-thread must not change the results"
        for _ in range(5):
            dev.cycle_device(True)
        dev.synchronize()
        dev.event_record(0)
        for _ in range(args.steps):
            dev.cycle_device(True)
        dev.event_record(1)
        ms = dev.event_elapsed_ms(0, 1) / args.steps
        dev.profile_enable(True)
        for _ in range(args.steps):
            dev.cycle_device(True)
        prof = dev.profile_report()
        dev.profile_enable(False)
        kern = {f"{r['kind']}{r['level']}": round(1e3 * r["total_ms"] / r["launches"], 1)
                for r in sorted(prof, key=lambda r: -r["total_ms"])[:14]}
        print(json.dumps({"rpt": rpt, "cycle_ms": round(ms, 4), "kernels_us": kern}), flush=True)
    dev.destroy()


if __name__ == "__main__":
    main()
