"""Host-side setup (product code, mirrors MGsetup.jl / SA-AMG.jl / GeometricTransferOperators.jl)."""
import warnings

import numpy as np
import pytest
import scipy.sparse as sp

import multigrid_jl_b200 as mg
from oracle import sa_literal


def test_1d_interp_cases():
    P, nc = mg.get1DFWInterp(9, False)                       # odd: tridiag(1/2,1,1/2)[:, ::2]
    assert nc == 5 and P.shape == (9, 5)
    np.testing.assert_array_equal(P.toarray()[:3, :2], [[1, 0], [0.5, 0.5], [0, 1]])
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        P, nc = mg.get1DFWInterp(8, True)                    # even + geometric: identity (+ warning)
        assert nc == 8 and (P != sp.identity(8)).nnz == 0 and len(w) == 1
    P, nc = mg.get1DFWInterp(2, False)
    assert nc == 2 and (P != sp.identity(2)).nnz == 0


@pytest.mark.parametrize("nn", [[9, 17], [9, 5, 17]])
def test_fw_interp_structure(nn):
    P, nc = mg.getFWInterp(nn)
    assert list(nc) == [(k + 1) // 2 for k in nn]
    assert P.nnz == np.prod([3 * c - 2 for c in nc])
    np.testing.assert_allclose(np.asarray(P.sum(axis=1)).ravel(), 1.0)
    assert set(np.unique(P.data)) <= {1.0, 0.5, 0.25, 0.125}


def test_level_sizes_match_survey_appendix_d():
    M = mg.getRegularMesh([0, 1, 0, 1], [128, 128])
    p = mg.getMGparam(np.float64, np.int64, 4, 8, 5, 1e-8, "Jac", 0.8, 2, 2, 'V')
    mg.MGsetup(mg.poisson_shifted(M), M, p, 1)
    assert [a.shape[0] for a in p.As] == [16641, 4225, 1089, 289]
    assert [a.nnz for a in p.As] == [82689, 37249, 9409, 2401]
    assert [P.nnz for P in p.Ps] == [37249, 9409, 2401]
    M = mg.getRegularMesh([0, 1, 0, 1, 0, 1], [32, 32, 32])
    p = mg.getMGparam(np.float64, np.int64, 3, 8, 5, 1e-8, "Jac", 0.8, 2, 2, 'V')
    mg.MGsetup(mg.poisson_shifted(M), M, p, 1)
    assert [a.shape[0] for a in p.As] == [35937, 4913, 729]
    assert [a.nnz for a in p.As] == [245025, 117649, 15625]
    for l in range(2):  # R = 2^-dim P', stored transposed (MGsetup.jl:56-60)
        assert abs(p.Rs[l] - 0.125 * p.Ps[l].T).max() == 0


@pytest.mark.parametrize("n", [[7, 5], [4, 6, 3]])
def test_direct_stencil_assembly_equals_GtG(n):
    rng = np.random.default_rng(0)
    M = mg.getRegularMesh([0, 1.0, 0, 2.0] + ([0, 0.5] if len(n) == 3 else []), n)
    G = mg.getNodalGradientMatrix(M)
    assert abs(mg.nodal_stencil_matrix(M) - G.T @ G).max() < 1e-12 * abs(G.T @ G).max()
    sigma = np.exp(rng.standard_normal(int(np.prod(n))))
    w = mg.edge_weights_from_cells(M, sigma)
    A = G.T @ sp.diags(np.concatenate(w)) @ G
    assert abs(mg.nodal_stencil_matrix(M, w) - A).max() < 1e-12 * abs(A).max()
    assert abs(mg.getNodalDivSigGradMatrix(M, sigma) - A).max() < 1e-12 * abs(A).max()
    B = mg.nodal_stencil_matrix(M, w)
    assert B.has_sorted_indices and (B.indices[:-1] < B.indices[1:])[np.diff(np.repeat(np.arange(B.shape[0]), np.diff(B.indptr))) == 0].all()


def test_relax_prec_formulas():
    rng = np.random.default_rng(3)
    A = sp.random(30, 30, density=0.2, random_state=4) + 1j * sp.random(30, 30, density=0.2, random_state=5)
    A = sp.csc_matrix(A + sp.diags(3.0 + rng.random(30) + 1j * rng.random(30)))
    AT = sp.csc_matrix(A.conj().T)
    d = mg.getRelaxPrec(AT, "Jac", 0.8)
    np.testing.assert_allclose(d, 0.8 / A.diagonal())                       # omega / a_ii
    d = mg.getRelaxPrec(AT, "SPAI", 1.0)
    colnorm2 = np.asarray(abs(A).power(2).sum(axis=0)).ravel()
    np.testing.assert_allclose(d, np.conj(A.diagonal()) / colnorm2)         # conj(a_ii)/||A e_i||^2


def test_coarsening_stops_when_no_dimension_can_coarsen():
    M = mg.getRegularMesh([0, 1, 0, 1], [12, 12])                           # 13 -> 7 -> 4 nodes (even): stop
    p = mg.getMGparam(np.float64, np.int64, 6, 8, 5, 1e-8, "Jac", 0.8, 2, 2, 'V')
    ctor = mg.getMultilevelOperatorConstructor([], lambda mesh: mg.poisson_shifted(mesh), [])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mg.MGsetup(ctor, M, p, 1)
    assert p.levels == 3 and [a.shape[0] for a in p.As] == [169, 49, 16]
    assert len(p.Ps) == len(p.Rs) == len(p.relaxPrecs) == 2
    assert all(a.nnz <= 5 * a.shape[0] for a in p.As)                       # rediscretised: 5-point everywhere


def test_transpose_and_replace_hierarchy():
    M = mg.getRegularMesh([0, 1, 0, 1], [16, 16])
    A = mg.poisson_shifted(M)
    p = mg.getMGparam(np.float64, np.int64, 3, 8, 5, 1e-8, "SPAI", 1.0, 1, 1, 'V')
    mg.MGsetup(A, M, p, 1)
    A2 = [a.copy() for a in p.As]
    mg.replaceMatrixInHierarchy(p, A * 2.0)
    assert abs(p.As[1] - 2.0 * A2[1]).max() < 1e-12 * abs(A2[1]).max()
    np.testing.assert_allclose(p.relaxPrecs[0], mg.getRelaxPrec(A * 2.0, "SPAI", 1.0))
    mg.transposeHierarchy(p)
    assert p.doTranspose == 1 and p.Ps[0].shape == A2[1].shape[:1] + A2[0].shape[:1]


# ---- SA-AMG: integer outputs bit-exact against the literal restatement -------------------------------
def _diffusion(n, seed):
    rng = np.random.default_rng(seed)
    M = mg.getRegularMesh([0, 1, 0, 1] + ([0, 1] if len(n) == 3 else []), n)
    sigma = np.exp(rng.standard_normal(int(np.prod(n))))
    w = mg.edge_weights_from_cells(M, sigma)
    A0 = mg.nodal_stencil_matrix(M, w, 0.0)
    return mg.nodal_stencil_matrix(M, w, 1e-8 * abs(A0).sum())


@pytest.mark.parametrize("n,seed", [([20, 20], 0), ([24, 17], 1), ([8, 7, 6], 2)])
def test_aggregation_bit_exact_vs_literal(n, seed):
    A = _diffusion(n, seed)
    S = mg.getStrengthMatrix(A, 0.4)
    S_lit = sa_literal.getStrengthMatrix(A, 0.4)
    assert np.array_equal(S.indptr, S_lit.indptr) and np.array_equal(S.indices, S_lit.indices)
    assert np.array_equal(S.data, S_lit.data)                               # bit-exact values too
    aggr = mg.neighborhoodAggregationNew(S)
    aggr_lit = sa_literal.neighborhoodAggregationNew(S_lit)
    assert np.array_equal(aggr, aggr_lit)
    P0, agg = mg.aggrArray2P(aggr)
    assert agg.min() == 1 and agg.max() == P0.shape[1] and P0.nnz == A.shape[0]
    roots = np.nonzero(aggr == np.arange(1, len(aggr) + 1))[0]
    assert np.array_equal(agg[roots], np.arange(1, len(roots) + 1))         # roots numbered in increasing index


def test_strength_matrix_properties():
    A = _diffusion([15, 15], 5)
    S = mg.getStrengthMatrix(A, 0.4)
    assert abs(S - S.T).max() == 0 and np.all(S.diagonal() == 2.0) and S.data.min() > 0


def test_sa_setup_hierarchy_and_opnorm_switch():
    A = _diffusion([50, 50], 7)
    p = mg.getMGparam(np.float64, np.int64, 4, 2, 5, 1e-8, "SPAI", 1.0, 1, 1, 'V', "Julia", 0.4)
    mg.SA_AMGsetup(A, p, True, 1)
    sizes = [a.shape[0] for a in p.As]
    assert sizes[0] == 2601 and all(a > b for a, b in zip(sizes, sizes[1:]))
    for l in range(p.levels - 1):
        assert p.Ps[l].shape == (sizes[l + 1], sizes[l]) and abs(p.Rs[l] - p.Ps[l].T).max() == 0
        assert p.aggregates[l].min() == 1 and p.aggregates[l].max() == sizes[l + 1]
    q = mg.getMGparam(np.float64, np.int64, 4, 2, 5, 1e-8, "SPAI", 1.0, 1, 1, 'V', "Julia", 0.4)
    mg.SA_AMGsetup(A, q, True, 1, opnorm=True)
    assert np.array_equal(q.aggregates[0], p.aggregates[0])                 # level-1 aggregation unaffected (A.6 item 5)
    with pytest.raises(RuntimeError):
        mg.SA_AMGsetup(A, q, False, 1)


def test_small_matrix_stops_sa_coarsening():
    A = _diffusion([9, 9], 1)                                               # 100 nodes <= 100 -> identity
    p = mg.getMGparam(np.float64, np.int64, 3, 2, 5, 1e-8, "SPAI", 1.0, 1, 1, 'V')
    mg.SA_AMGsetup(A, p, True, 1)
    assert p.levels == 1 and len(p.Ps) == 0
