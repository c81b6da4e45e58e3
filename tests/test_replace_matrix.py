"""replaceMatrixInHierarchy (MGsetup.jl:226-270) on the device: mgb200_replace_matrix (csrc/galerkin.cuh) against the host
restatement of the same function (multigrid.jl_b200/mgsetup.py, scipy products) and, for the solve that follows, against
the CPU oracle on the host-recomputed hierarchy.  Tolerances: Galerkin values 1e-13 relative to the largest entry of the
level (the two products sum in different orders), relaxation weights 1e-14, per-cycle residual norms 1e-10."""
import copy

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import make_problem


def _second_matrix(kind, M, seed=7):
    """A matrix with the sparsity of make_problem(kind)'s and other values."""
    import multigrid_jl_b200 as mg
    rng = np.random.default_rng(seed)
    if kind == "poisson":
        return mg.poisson_shifted(M, 3e-2)
    if kind == "diffusion":
        sigma = np.exp(rng.standard_normal(int(np.prod(M.n))))
        w = mg.edge_weights_from_cells(M, sigma)
        A0 = mg.nodal_stencil_matrix(M, w, 0.0)
        return mg.nodal_stencil_matrix(M, w, 1e-5 * abs(A0).sum(axis=0).max())
    if kind == "helmholtz":
        kappa = 2 * np.pi / (10 * M.h[0]) * 0.25
        return mg.helmholtz_shifted(M, kappa ** 2, 0.3)
    raise ValueError(kind)


def _host_twin(p, AT2):
    """The same hierarchy redone on the host (reference order), without touching p's device."""
    import multigrid_jl_b200 as mg
    q = copy.copy(p)
    q.As, q.Ps, q.Rs, q.relaxPrecs = list(p.As), list(p.Ps), list(p.Rs), list(p.relaxPrecs)
    q.device = None
    q._mixed_device = None
    mg.replaceMatrixInHierarchy(q, AT2, device=False)
    return q


@pytest.mark.gpu
@pytest.mark.parametrize("kind,n,levels,relax", [("poisson", [24, 20, 16], 3, "Jac"), ("poisson", [64, 48], 4, "SPAI"),
                                                 ("diffusion", [16, 16, 16], 3, "Jac"), ("diffusion", [20, 12, 16], 3, "SPAI"),
                                                 ("helmholtz", [16, 16, 16], 3, "Jac"), ("helmholtz", [40, 32], 3, "SPAI")])
def test_replace_matrix_on_the_device(kind, n, levels, relax):
    import multigrid_jl_b200 as mg
    from oracle import cycle as oc
    A, AT, M, p, b = make_problem(kind, n, levels, relax=relax, omega=0.8 if relax == "Jac" else 1.0, maxit=4)
    dev = mg.uploadHierarchy(p)
    x, _, it0 = mg.solveMG(p, b, np.zeros_like(b))          # the hierarchy is in use before it is replaced
    A2 = _second_matrix(kind, M)
    AT2 = A2.conj().T.tocsc() if p.VAL == np.complex128 else sp.csc_matrix(A2)
    AT2.sort_indices()
    q = _host_twin(p, AT2)
    mg.replaceMatrixInHierarchy(p, AT2)
    assert p.device is dev, "the device path must keep the resident hierarchy"
    for l in range(p.levels):
        ref, got = sp.csc_matrix(q.As[l]), sp.csc_matrix(p.As[l])
        ref.sort_indices()
        assert np.array_equal(ref.indptr, got.indptr) and np.array_equal(ref.indices, got.indices)
        assert np.abs(got.data - ref.data).max() <= 1e-13 * np.abs(ref.data).max(), l
    for l in range(p.levels - 1):
        assert np.abs(p.relaxPrecs[l] - q.relaxPrecs[l]).max() <= 1e-14 * np.abs(q.relaxPrecs[l]).max(), l
    # solve with the replaced hierarchy against the oracle on the host-recomputed one
    b2 = A2 @ (np.random.default_rng(3).random(A2.shape[0]).astype(p.VAL))
    b2 = np.asfortranarray((b2 / np.linalg.norm(b2)).astype(p.VAL))
    o = oc.OracleMG(q)
    x_ref, it_ref, res_ref = oc.solveMG(o, b2, np.zeros_like(b2))
    x, _, it = mg.solveMG(p, b2, np.zeros_like(b2))
    assert it == it_ref
    np.testing.assert_allclose(p.last_resvec, res_ref, rtol=1e-10, atol=0)
    assert np.linalg.norm(x - x_ref) <= 1e-9 * np.linalg.norm(x_ref)


@pytest.mark.gpu
def test_replace_matrix_keeps_the_dictionary_of_constant_coefficient_levels():
    """Galerkin levels of a constant-coefficient operator keep their stencil dictionary (every row still equals its
    pattern's representative row); a variable-coefficient replacement drops it and runs the CSR kernels."""
    import multigrid_jl_b200 as mg
    A, AT, M, p, b = make_problem("poisson", [32, 32, 32], 3, maxit=3)
    dev = mg.uploadHierarchy(p)
    before = dev.pattern_info(2, 0)
    assert before["in_use"]
    mg.replaceMatrixInHierarchy(p, sp.csc_matrix(_second_matrix("poisson", M)))
    after = dev.pattern_info(2, 0)
    assert after["in_use"] and after["patterns"] == before["patterns"] and after["d_folded"] == before["d_folded"]
    mg.replaceMatrixInHierarchy(p, sp.csc_matrix(_second_matrix("diffusion", M)))
    assert p.device is dev
    assert not dev.pattern_info(2, 0)["in_use"]
    x, _, it = mg.solveMG(p, b, np.zeros_like(b))
    assert p.last_resvec[-1] < 1e-2 * p.last_resvec[0]


@pytest.mark.gpu
def test_replace_matrix_with_another_sparsity_falls_back_to_the_host():
    import multigrid_jl_b200 as mg
    A, AT, M, p, b = make_problem("poisson", [16, 16, 16], 3, maxit=3)
    dev = mg.uploadHierarchy(p)
    G = mg.getNodalGradientMatrix(M)
    L = (G.T @ G).tocsr()
    A2 = sp.csc_matrix(L @ L * 1e-4 + L + 1e-2 * sp.identity(L.shape[0]))        # wider stencil
    A2.sort_indices()
    assert A2.nnz != p.As[0].nnz
    assert dev.replace_matrix(A2, "Jac", 0.8) is False
    mg.replaceMatrixInHierarchy(p, A2)
    assert p.device is None          # dropped for re-upload, as the host path always did
    x, _, it = mg.solveMG(p, b, np.zeros_like(b))
    assert p.device is not None and p.last_resvec[-1] < p.last_resvec[0]


@pytest.mark.gpu
def test_replace_matrix_sa_amg():
    """SA-AMG hierarchy (general CSR P with several entries per row, no dictionaries): device products against scipy."""
    import multigrid_jl_b200 as mg
    n = [16, 16, 12]
    M = mg.getRegularMesh([0, 1, 0, 1, 0, 1], n)
    rng = np.random.default_rng(0)

    def mat(seed):
        r = np.random.default_rng(seed)
        w = mg.edge_weights_from_cells(M, np.exp(r.standard_normal(int(np.prod(n)))))
        A0 = mg.nodal_stencil_matrix(M, w, 0.0)
        return sp.csc_matrix(mg.nodal_stencil_matrix(M, w, 1e-6 * abs(A0).sum(axis=0).max()))
    A1, A2 = mat(1), mat(2)
    p = mg.getMGparam(np.float64, np.int64, 3, 8, 4, 1e-12, "SPAI", 1.0, 2, 2, 'W')
    mg.SA_AMGsetup(A1, p, True, 1)
    dev = mg.uploadHierarchy(p)
    q = _host_twin(p, A2)
    mg.replaceMatrixInHierarchy(p, A2)
    assert p.device is dev
    for l in range(p.levels):
        ref, got = sp.csc_matrix(q.As[l]), sp.csc_matrix(p.As[l])
        ref.sort_indices()
        assert np.abs(got.data - ref.data).max() <= 1e-13 * np.abs(ref.data).max(), l
    b = A2 @ rng.random(A2.shape[0])
    b /= np.linalg.norm(b)
    from oracle import cycle as oc
    x_ref, it_ref, res_ref = oc.solveMG(oc.OracleMG(q), b, np.zeros_like(b))
    x, _, it = mg.solveMG(p, b, np.zeros_like(b))
    np.testing.assert_allclose(p.last_resvec, res_ref, rtol=1e-10, atol=0)


def test_replace_matrix_symbols_exported():
    from multigrid_jl_b200 import device
    L = device.lib()
    for name in ("mgb200_replace_matrix", "mgb200_download_values", "mgb200_download_relax_prec"):
        assert hasattr(L, name)
