"""Committed golden fixtures (tests/golden/golden_v1.npz, made by tests/golden/make_golden.py from
the CPU oracle - see that script's header for what they do and do not pin)."""
import os

import numpy as np
import pytest

from conftest import make_problem

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_v1.npz"))
CASES = {"poisson2d": ("poisson", [16, 16], 3), "poisson3d": ("poisson", [8, 8, 8], 3),
         "helmholtz2d": ("helmholtz", [16, 16], 3)}


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("cyc", list("VWFK"))
def test_oracle_reproduces_golden(name, cyc):
    from oracle import cycle as oc
    kind, n, levels = CASES[name]
    A, AT, M, p, b = make_problem(kind, n, levels, cycle=cyc, maxit=5)
    np.testing.assert_array_equal(b, G[f"{name}_{cyc}_b"])          # same seeded input
    x, it, res = oc.solveMG(oc.OracleMG(p), b, np.zeros_like(b))
    np.testing.assert_allclose(res, G[f"{name}_{cyc}_res"], rtol=1e-10)  # OpenMP reductions are not bit-reproducible
    np.testing.assert_allclose(x, G[f"{name}_{cyc}_x"], rtol=1e-8, atol=1e-13)


def test_sa_aggregation_golden():
    import multigrid_jl_b200 as mg
    rng = np.random.default_rng(11)
    Mm = mg.getRegularMesh([0, 1, 0, 1], [20, 20])
    w = mg.edge_weights_from_cells(Mm, np.exp(rng.standard_normal(400)))
    A0 = mg.nodal_stencil_matrix(Mm, w, 0.0)
    Asa = mg.nodal_stencil_matrix(Mm, w, 1e-8 * abs(A0).sum())
    S = mg.getStrengthMatrix(Asa, 0.4)
    assert np.array_equal(S.indptr, G["sa_S_indptr_20x20"]) and np.array_equal(S.indices, G["sa_S_indices_20x20"])
    assert np.array_equal(mg.neighborhoodAggregationNew(S), G["sa_aggr_20x20"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("cyc", list("VWFK"))
def test_gpu_reproduces_golden(name, cyc):
    import multigrid_jl_b200 as mg
    kind, n, levels = CASES[name]
    A, AT, M, p, b = make_problem(kind, n, levels, cycle=cyc, maxit=5)
    x = np.zeros_like(b)
    mg.solveMG(p, G[f"{name}_{cyc}_b"], x)
    np.testing.assert_allclose(p.last_resvec, G[f"{name}_{cyc}_res"], rtol=1e-10)
    np.testing.assert_allclose(x, G[f"{name}_{cyc}_x"], rtol=1e-8, atol=1e-13)


@pytest.mark.gpu
def test_gpu_krylov_golden():
    import multigrid_jl_b200 as mg
    A, AT, M, p, b = make_problem("poisson", [16, 16], 3, maxit=30, tol=1e-8)
    x = np.zeros_like(b)
    x, _, it = mg.solveCG_MG(AT, p, b, x)
    assert [it, p.last_flag] == list(G["cg_poisson2d_iter"])
    np.testing.assert_allclose(p.last_resvec, G["cg_poisson2d_resvec"], rtol=1e-8)
    A, AT, M, p, b = make_problem("helmholtz", [16, 16], 3, maxit=10, tol=1e-8)
    x = np.zeros_like(b)
    x, _, it, res = mg.solveGMRES_MG(AT, p, b, x, True, 5)
    assert [it, p.last_flag] == list(G["fgmres_helmholtz2d_iter"])
    np.testing.assert_allclose(res, G["fgmres_helmholtz2d_resvec"], rtol=1e-7)
