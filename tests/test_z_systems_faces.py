"""Staggered-grid (faces) transfer operators (Systems.jl; multigrid.jl_b200/systems.py) and the geometric setup with
``transferOperatorType`` "SystemsFacesLinear" / "SystemsFacesMixedLinear" (MGsetup.jl:49-75).  CPU: the 1-D builders
against the stencils the reference's code produces, block structure, row sums, Galerkin hierarchy, convergence of the
oracle's cycle.  GPU: parity of the device cycle with the oracle on such a hierarchy (general-CSR path).
(Sorts last like test_z_classical_amg.py: the GPU case has not run on a GPU yet.)"""
import numpy as np
import pytest
import scipy.sparse as sp


def _lap_nodes(n, h):
    G = sp.diags([-np.ones(n), np.ones(n)], [0, 1], shape=(n, n + 1)) / h
    return (G.T @ G).tocsc()


def _lap_cells(n, h):
    G = sp.diags([-np.ones(n - 1), np.ones(n - 1)], [0, 1], shape=(n - 1, n)) / h
    return (G.T @ G).tocsc()


def faces_operator(n, shift=1e-3, cells_block=False):
    """Block-diagonal vector Laplacian on a staggered grid: per face block nodes in its own direction, cells in the
    others (Neumann), plus a cell-centred block; shifted to be definite.  Unknown ordering as in systems.py."""
    dim = len(n)
    h = [1.0 / k for k in n]
    blocks = []
    for j in range(dim + (1 if cells_block else 0)):
        ops = [_lap_nodes(n[k], h[k]) if k == j else _lap_cells(n[k], h[k]) for k in range(dim)]
        sizes = [o.shape[0] for o in ops]
        L = sp.csc_matrix((int(np.prod(sizes)),) * 2)
        for k in range(dim):
            mats = [ops[q] if q == k else sp.identity(sizes[q]) for q in range(dim)]
            K = mats[0]
            for q in range(1, dim):
                K = sp.kron(mats[q], K, format="csc")
            L = L + K
        blocks.append(L)
    A = sp.block_diag(blocks, format="csc")
    A = sp.csc_matrix(A + shift * abs(A).sum(axis=0).max() * sp.identity(A.shape[0], format="csc"))
    A.sort_indices()
    return A


def _setup(n, typ, levels=3, cycle='V'):
    import multigrid_jl_b200 as mg
    A = faces_operator(n, 1e-3, typ.endswith("MixedLinear"))
    p = mg.getMGparam(np.float64, np.int64, levels, 8, 6, 1e-12, "Jac", 0.7, 2, 2, cycle, "NoMUMPS", 0.4, 0.0, typ)
    mg.MGsetup(A, mg.getRegularMesh([0.0, 1.0] * len(n), n), p, 1)
    b = A @ np.random.default_rng(0).random(A.shape[0])
    return A, p, b / np.linalg.norm(b)


def test_1d_builders():
    from multigrid_jl_b200 import systems as S
    P, nc = S.get1DProlongationCellCentered(8)
    assert nc == 4 and P.shape == (8, 4)
    np.testing.assert_array_equal(P.toarray()[:5, :3], [[1, 0, 0], [.75, .25, 0], [.25, .75, 0], [0, .75, .25], [0, .25, .75]])
    np.testing.assert_array_equal(P.toarray()[-2:, -2:], [[.25, .75], [0, 1]])
    R, _ = S.get1DRestrictionCells(8)
    np.testing.assert_array_equal(R.toarray()[1], [0, 0, 1, 1, 0, 0, 0, 0])
    Rn, _ = S.get1DNodeFullWeightRestriction(8)
    assert Rn.shape == (5, 9)
    np.testing.assert_array_equal(Rn.toarray()[0, :3], [1, .5, 0])
    np.testing.assert_array_equal(Rn.toarray()[2, 2:7], [0, .5, 1, .5, 0])
    Pn, _ = S.get1DProlongationNodes(8)
    np.testing.assert_array_equal(Pn.toarray()[:3, :2], [[1, 0], [.5, .5], [0, 1]])
    Ri, _ = S.get1DNodeInjection(8)
    assert Ri.shape == (5, 9) and Ri.nnz == 5 and np.all(Ri.tocsr().indices == np.arange(0, 9, 2))
    # fewer than 8 cells: that dimension is not coarsened (identity), odd sizes are an error
    for f, size in ((S.get1DProlongationNodes, 7), (S.get1DProlongationCellCentered, 6), (S.get1DRestrictionCells, 6)):
        M, nc = f(6)
        assert nc == 6 and M.shape == (size, size) and (M - sp.identity(size)).nnz == 0
    with pytest.raises(ValueError):
        S.get1DRestrictionCells(9)


@pytest.mark.parametrize("n,cells", [([8, 12], False), ([8, 12, 16], False), ([16, 8], True), ([8, 8, 8], True)])
def test_block_operators(n, cells):
    import multigrid_jl_b200 as mg
    P, R, nc = mg.getLinearOperatorsSystemsFaces(n, cells)
    assert list(nc) == [k // 2 for k in n]
    assert P.shape == (mg.faces_size(n, cells), mg.faces_size(nc, cells)) and R.shape == P.shape[::-1]
    # block diagonal: no coupling between components
    off_f, off_c = 0, 0
    for j in range(len(n) + (1 if cells else 0)):
        nf = int(np.prod([n[k] + (1 if k == j else 0) for k in range(len(n))]))
        ncj = int(np.prod([nc[k] + (1 if k == j else 0) for k in range(len(n))]))
        sub = P[off_f:off_f + nf, :]
        assert sub[:, :off_c].nnz == 0 and sub[:, off_c + ncj:].nnz == 0
        off_f += nf
        off_c += ncj
    # prolongation reproduces constants; restriction x 2^-dim (MGsetup.jl:70-74) averages away from the boundary nodes
    np.testing.assert_allclose(P @ np.ones(P.shape[1]), 1.0, atol=1e-15)
    rs = np.asarray(R.sum(axis=1)).ravel() * 0.5 ** len(n)
    assert rs.max() == 1.0 and rs.min() == 0.75 and np.median(rs) == 1.0
    Rinj = mg.getInjectionOperatorsSystemsFaces(n, cells)
    assert Rinj.shape == R.shape


@pytest.mark.parametrize("n,typ", [([16, 16], "SystemsFacesLinear"), ([16, 8, 16], "SystemsFacesLinear"),
                                   ([16, 16], "SystemsFacesMixedLinear")])
def test_setup_and_oracle_convergence(n, typ):
    import multigrid_jl_b200 as mg
    from oracle import cycle as oc
    A, p, b = _setup(n, typ)
    cb = typ.endswith("MixedLinear")
    assert [a.shape[0] for a in p.As] == [mg.faces_size(m.n, cb) for m in p.Meshes]
    for l in range(p.levels - 1):
        G = p.Ps[l] @ p.As[l] @ p.Rs[l]
        assert abs(G - p.As[l + 1]).max() <= 1e-13 * abs(G).max()
    x, it, res = oc.solveMG(oc.OracleMG(p), b, np.zeros_like(b))
    assert res[-1] < (1e-4 if len(n) == 2 else 5e-2) * res[0] and np.all(res[1:] < res[:-1])


@pytest.mark.gpu
@pytest.mark.parametrize("n,typ,cycle", [([16, 16, 8], "SystemsFacesLinear", 'V'), ([32, 16], "SystemsFacesMixedLinear", 'W')])
def test_device_cycle_on_a_faces_hierarchy(n, typ, cycle):
    import multigrid_jl_b200 as mg
    from oracle import cycle as oc
    A, p, b = _setup(n, typ, cycle=cycle)
    x_ref, it_ref, res_ref = oc.solveMG(oc.OracleMG(p), b, np.zeros_like(b))
    x, _, it = mg.solveMG(p, b, np.zeros_like(b))
    assert it == it_ref
    np.testing.assert_allclose(p.last_resvec, res_ref, rtol=1e-10, atol=0)
    assert np.linalg.norm(x - x_ref) <= 1e-9 * np.linalg.norm(x_ref)
