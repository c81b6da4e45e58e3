"""The reference's own integration tests re-expressed on SEEDED inputs and run through the CPU
oracle: the only acceptance floor the reference's test-suite provides for this path (residual
thresholds).  The GPU twin of each case is in test_gpu_parity.py."""
import numpy as np

import multigrid_jl_b200 as mg
from oracle import cycle as oc


def _poisson(n, shift=1e-4):
    dom = [0.0, 1.0] * len(n)
    M = mg.getRegularMesh(dom, n)
    return M, mg.poisson_shifted(M, shift)


def test_testGMGRAPforPoisson_2d():
    """test/Multigrid/testGMGRAPforPoisson.jl:8-40: 128x128, 2 RHS, Jac-GMRES 0.75, V(1,1), 4 levels,
    5 cycles -> norm(A x - b) < 0.005 with ||b|| = 1."""
    M, A = _poisson([128, 128])
    p = mg.getMGparam(np.float64, np.int64, 4, 8, 5, 1e-10, "Jac-GMRES", 0.75, 1, 1, 'V', "NoMUMPS", 0.5, 0.0)
    mg.MGsetup(A, M, p, 2)
    rng = np.random.default_rng(0)
    b = np.asfortranarray(A @ rng.random((A.shape[0], 2)))
    b /= np.linalg.norm(b)
    x, it, res = oc.solveMG(oc.OracleMG(p), b, np.zeros_like(b))
    assert np.linalg.norm(A @ x - b) < 0.005


def test_testGMGRAPforPoisson_3d():
    """:59-78: n = [32,32,16] -> < 0.01"""
    M, A = _poisson([32, 32, 16])
    p = mg.getMGparam(np.float64, np.int64, 4, 8, 5, 1e-10, "Jac-GMRES", 0.75, 1, 1, 'V', "NoMUMPS", 0.5, 0.0)
    mg.MGsetup(A, M, p, 2)
    rng = np.random.default_rng(1)
    b = np.asfortranarray(A @ rng.random((A.shape[0], 2)))
    b /= np.linalg.norm(b)
    x, it, res = oc.solveMG(oc.OracleMG(p), b, np.zeros_like(b))
    assert np.linalg.norm(A @ x - b) < 0.01


def test_testGMG_jacobi():
    """test/Multigrid/testGMG.jl:21-37,65-68: 128x128, Jac 0.8, V(1,1), 4 levels, 5 cycles, Galerkin
    from the matrix -> < 0.005 (the Neumann operator is made definite with the usual shift)."""
    M, A = _poisson([128, 128])
    p = mg.getMGparam(np.float64, np.int64, 4, 8, 5, 1e-2, "Jac", 0.8, 1, 1, 'V')
    mg.MGsetup(A, M, p, 1)
    rng = np.random.default_rng(2)
    b = A @ rng.random(A.shape[0])
    b /= np.linalg.norm(b)
    x, it, res = oc.solveMG(oc.OracleMG(p), b, np.zeros_like(b))
    assert np.linalg.norm(A @ x - b) < 0.005


def test_testSAforDivSigGrad():
    """test/Multigrid/testSAforDivSigGrad.jl:9-44: 50x50, sigma = exp(randn), SPAI 1.0, V(1,1), 3 levels,
    3 RHS: solveMG < 0.01, then CG < 0.005."""
    rng = np.random.default_rng(3)
    M = mg.getRegularMesh([0, 1, 0, 1], [50, 50])
    w = mg.edge_weights_from_cells(M, np.exp(rng.standard_normal(2500)))
    A0 = mg.nodal_stencil_matrix(M, w, 0.0)
    A = mg.nodal_stencil_matrix(M, w, 1e-8 * abs(A0).sum())
    p = mg.getMGparam(np.float64, np.int64, 3, 2, 5, 1e-4, "SPAI", 1.0, 1, 1, 'V', "Julia")
    mg.SA_AMGsetup(A, p, True, 3)
    b = np.asfortranarray(A @ rng.random((A.shape[0], 3)))
    b /= np.linalg.norm(b)
    o = oc.OracleMG(p)
    x, it, res = oc.solveMG(o, b, np.zeros_like(b))
    assert np.linalg.norm(A @ x - b) < 0.01
    x, it, flag, resv = oc.solveCG_MG(A, o, b, np.zeros_like(b))
    assert np.linalg.norm(A @ x - b) < 0.005
