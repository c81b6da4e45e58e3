"""Generates tests/golden/golden_v1.npz.

These vectors are produced by THIS repository's CPU oracle (oracle/), not by the reference
(which cannot run here: no julia / gfortran).  They pin the oracle and the
GPU path against regressions and give both the same committed inputs; they do not pin the oracle
against the reference.  Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import make_problem  # noqa: E402
from oracle import cycle as oc  # noqa: E402
import multigrid_jl_b200 as mg  # noqa: E402

CASES = {
    "poisson2d": dict(kind="poisson", n=[16, 16], levels=3),
    "poisson3d": dict(kind="poisson", n=[8, 8, 8], levels=3),
    "helmholtz2d": dict(kind="helmholtz", n=[16, 16], levels=3),
}


def main():
    out = {}
    for name, c in CASES.items():
        for cyc in "VWFK":
            A, AT, M, p, b = make_problem(c["kind"], c["n"], c["levels"], cycle=cyc, maxit=5)
            o = oc.OracleMG(p)
            x, it, res = oc.solveMG(o, b, np.zeros_like(b))
            out[f"{name}_{cyc}_b"] = b
            out[f"{name}_{cyc}_res"] = res
            out[f"{name}_{cyc}_x"] = x
    A, AT, M, p, b = make_problem("poisson", [16, 16], 3, maxit=30, tol=1e-8)
    o = oc.OracleMG(p)
    x, it, flag, resv = oc.solveCG_MG(AT, o, b, np.zeros_like(b))
    out["cg_poisson2d_resvec"] = resv
    out["cg_poisson2d_iter"] = np.array([it, flag])
    A, AT, M, p, b = make_problem("helmholtz", [16, 16], 3, maxit=10, tol=1e-8)
    o = oc.OracleMG(p)
    x, it, flag, resv = oc.solveGMRES_MG(AT, o, b, np.zeros_like(b), True, 5)
    out["fgmres_helmholtz2d_resvec"] = resv
    out["fgmres_helmholtz2d_iter"] = np.array([it, flag])
    # SA-AMG integer maps of a seeded 20x20 diffusion problem
    rng = np.random.default_rng(11)
    Mm = mg.getRegularMesh([0, 1, 0, 1], [20, 20])
    w = mg.edge_weights_from_cells(Mm, np.exp(rng.standard_normal(400)))
    A0 = mg.nodal_stencil_matrix(Mm, w, 0.0)
    Asa = mg.nodal_stencil_matrix(Mm, w, 1e-8 * abs(A0).sum())
    S = mg.getStrengthMatrix(Asa, 0.4)
    out["sa_aggr_20x20"] = mg.neighborhoodAggregationNew(S)
    out["sa_S_indptr_20x20"] = S.indptr.astype(np.int64)
    out["sa_S_indices_20x20"] = S.indices.astype(np.int64)
    np.savez_compressed(os.path.join(HERE, "golden_v1.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
