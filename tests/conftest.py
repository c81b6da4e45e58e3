import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """The oracle's C kernels are built on demand; the CUDA library is built here when nvcc is
    available and the .so is missing (on the GPU box the prebuilt file travels with the repo)."""
    import __graft_entry__ as g
    g.build_oracle()
    if not os.path.exists(g.LIB):
        g.build_cuda()
    yield


def make_problem(kind, n, levels, relax="Jac", omega=0.8, pre=2, post=2, cycle='V', nrhs=1, seed=0,
                 maxit=8, tol=1e-12):
    """Seeded synthetic problems shared by the parity tests."""
    import multigrid_jl_b200 as mg
    rng = np.random.default_rng(seed)
    dim = len(n)
    dom = [0.0, 1.0] * dim
    M = mg.getRegularMesh(dom, n)
    if kind == "poisson":
        A = mg.poisson_shifted(M, 1e-4)
        VAL = np.float64
    elif kind == "diffusion":
        sigma = np.exp(rng.standard_normal(int(np.prod(n))))
        w = mg.edge_weights_from_cells(M, sigma)
        A0 = mg.nodal_stencil_matrix(M, w, 0.0)
        A = mg.nodal_stencil_matrix(M, w, 1e-6 * abs(A0).sum(axis=0).max())
        VAL = np.float64
    elif kind == "helmholtz":
        h = M.h[0]
        kappa = 2 * np.pi / (10 * h) * 0.35  # mild: the V-cycle must converge for a parity run
        A = mg.helmholtz_shifted(M, kappa ** 2, 0.5)
        VAL = np.complex128
    else:
        raise ValueError(kind)
    p = mg.getMGparam(VAL, np.int64, levels, 8, maxit, tol, relax, omega, pre, post, cycle)
    AT = A.conj().T.tocsc() if VAL == np.complex128 else A
    AT.sort_indices()
    if kind == "sa":
        raise ValueError
    mg.MGsetup(AT, M, p, nrhs)
    N = A.shape[0]
    shape = (N,) if nrhs == 1 else (N, nrhs)
    u = rng.random(shape)
    if VAL == np.complex128:
        u = u + 1j * rng.random(shape)
    b = np.asfortranarray(A @ u)
    b = b / np.linalg.norm(b)
    return A, AT, M, p, np.asfortranarray(b.astype(VAL))
