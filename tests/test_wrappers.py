"""jInv AbstractSolver plug-ins (MGWrapper.jl, SAAMGWrapper.jl): the reference's own wrapper tests
(test/Multigrid/testLinSolveMGWrapper.jl:20-40, testSAforDivSigGrad.jl:95-116) re-run on seeded inputs
through the device path."""
import numpy as np
import pytest
import scipy.sparse as sp


def _problem(n, seed=0):
    import multigrid_jl_b200 as mg
    M = mg.getRegularMesh([0.0, 1.0] * len(n), n)
    A = mg.poisson_shifted(M, 1e-4)
    rng = np.random.default_rng(seed)
    B = A @ rng.random(A.shape[0])
    return mg, M, A, B


def test_wrapper_defaults_and_zero_rhs_need_no_device():
    """B = 0 returns X = 0 before any setup (MGWrapper.jl:38-41); counters start at zero."""
    mg, M, A, B = _problem([16, 16])
    MG = mg.getMGparam(np.float64, np.int64, 3, 8, 15, 1e-2, "SPAI", 1.0, 2, 2, 'V', "Julia")
    s = mg.getMGsolver(MG, M, 1, "GMRES", out=-1)
    assert s.tol == 1e-2 and s.nIter == 0 and s.Krylov == "GMRES" and MG.Meshes == [M]
    X = np.ones(A.shape[0])
    X, s2 = mg.solveLinearSystem(A, np.zeros(A.shape[0]), X, s)
    assert s2 is s and not X.any() and not mg.hierarchyExists(MG)
    sa = mg.getSA_AMGsolver(mg.getMGparam(np.float64, np.int64, 3, 8, 15, 1e-2, "SPAI", 1.0, 1, 1, 'V'), "PCG")
    assert sa.sym == 1 and sa.Krylov == "PCG"
    c = mg.wrappers.copySolver(s)
    assert c.MG is not MG and c.Krylov == s.Krylov and c.nIter == 0 and not mg.hierarchyExists(c.MG)


@pytest.mark.gpu
@pytest.mark.parametrize("krylov", ["GMRES", "PCG", "BiCGSTAB", "MG"])
def test_mgsolver_reference_wrapper_test(krylov):
    """testLinSolveMGWrapper.jl:20-40: SPAI V(2,2), tol 1e-2, lazy setup, residual below the tolerance."""
    mg, M, A, B = _problem([48, 48])
    MG = mg.getMGparam(np.float64, np.int64, 4, 8, 15, 1e-2, "SPAI", 1.0, 2, 2, 'V', "Julia")
    s = mg.getMGsolver(MG, M, 1, krylov, out=-1)
    X = np.zeros_like(B)
    X, s = mg.solveLinearSystem(A, B, X, s)
    assert mg.hierarchyExists(MG) and s.nIter > 0 and s.timeSetup > 0 and s.timeSolve > 0
    assert np.linalg.norm(A @ X - B) / np.linalg.norm(B) < s.tol
    it1 = s.nIter
    X2, s = mg.solveLinearSystem(A, B, np.zeros_like(B), s)     # the hierarchy is reused
    assert s.nIter == 2 * it1
    mg.wrappers.clear(s)
    assert not mg.hierarchyExists(MG) and s.doClear == 0


@pytest.mark.gpu
def test_mgsolver_transposed_nonsymmetric_solve():
    """sym = 0: A is transposed for the setup (MGWrapper.jl:55-59) and doTranspose = 1 solves with A^T through
    transposeHierarchy (:65-67)."""
    mg, M, A, B = _problem([32, 32])
    G = sp.diags([np.full(A.shape[0] - 1, 0.4 * A.diagonal().mean() / 4)], [1], format="csc")
    An = sp.csc_matrix(A + G)                                   # nonsymmetric perturbation
    MG = mg.getMGparam(np.float64, np.int64, 3, 8, 40, 1e-6, "Jac", 0.8, 2, 2, 'V')
    s = mg.getMGsolver(MG, M, 0, "GMRES", out=-1)
    X = np.zeros_like(B)
    X, s = mg.solveLinearSystem(An, B, X, s, 0)
    assert np.linalg.norm(An @ X - B) / np.linalg.norm(B) < 1e-5
    Xt = np.zeros_like(B)
    Xt, s = mg.solveLinearSystem(An, B, Xt, s, 1)
    assert MG.doTranspose == 1
    assert np.linalg.norm(An.T @ Xt - B) / np.linalg.norm(B) < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("krylov", ["BiCGSTAB", "PCG"])
def test_sa_amgsolver(krylov):
    """testSAforDivSigGrad.jl:95-116 in small: SA-AMG + SPAI as a jInv solver."""
    import multigrid_jl_b200 as mg
    rng = np.random.default_rng(1)
    n = [24, 24, 12]
    M = mg.getRegularMesh([0.0, 1.0] * 3, n)
    w = mg.edge_weights_from_cells(M, np.exp(rng.standard_normal(int(np.prod(n)))))
    A0 = mg.nodal_stencil_matrix(M, w, 0.0)
    A = mg.nodal_stencil_matrix(M, w, 1e-6 * abs(A0).sum(axis=0).max())
    B = A @ rng.random(A.shape[0])
    MG = mg.getMGparam(np.float64, np.int64, 3, 8, 30, 1e-6, "SPAI", 1.0, 1, 1, 'V', "Julia", 0.4)
    s = mg.getSA_AMGsolver(MG, krylov, sym=1, out=-1)
    X = np.zeros_like(B)
    X, s = mg.solveLinearSystem(A, B, X, s)
    assert np.linalg.norm(A @ X - B) / np.linalg.norm(B) < 1e-5 and s.nIter > 0
