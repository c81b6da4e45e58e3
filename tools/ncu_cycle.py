#!/usr/bin/env python
"""One V(2,2) cycle of the bench workload inside a cudaProfilerStart/Stop bracket, for `ncu --profile-from-start off`.
Kernel options come from the environment (MGB200_*)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import multigrid_jl_b200 as mg  # noqa: E402
from bench import build_problem  # noqa: E402
from multigrid_jl_b200.device import lib  # noqa: E402

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 256
levels = int(sys.argv[2]) if len(sys.argv) > 2 else 6
A, M, p, b = build_problem(cells, levels)
dev = mg.DeviceHierarchy(p, device=0)
dev.set_option("graphs", 0)          # eager launches: ncu sees every kernel with its own launch configuration
x = np.zeros_like(b)
dev.solveMG(b, x, 0.0, 1)
for _ in range(2):
    dev.cycle_device(True)
dev.synchronize()
lib().mgb200_profiler_start()
dev.cycle_device(True)
dev.synchronize()
lib().mgb200_profiler_stop()
dev.destroy()
