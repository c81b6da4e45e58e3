"""The C-ABI library loads on a CPU-only box, exports every symbol include/mgb200.h declares, and
fails loudly (no CPU fallback) when asked to compute without a GPU.  No compute calls here."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "mgb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mgb200_[a-zA-Z0-9_]+)\s*\(", src)))


def test_exports_every_declared_symbol():
    from multigrid_jl_b200 import device
    L = device.lib()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/mgb200.h but not exported"
    assert L.mgb200_version() >= 100


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    import multigrid_jl_b200 as mg
    M = mg.getRegularMesh([0, 1, 0, 1], [8, 8])
    p = mg.getMGparam(np.float64, np.int64, 2, 8, 5, 1e-8, "Jac", 0.8, 2, 2, 'V')
    mg.MGsetup(mg.poisson_shifted(M), M, p, 1)
    b = np.ones(81)
    with pytest.raises(mg.MGB200Error, match="no CPU fallback"):
        mg.solveMG(p, b, np.zeros(81))
    with pytest.raises(mg.MGB200Error):
        mg.solveCG_MG(p.As[0], p, b, np.zeros(81))


def test_null_handle_is_an_error_not_a_crash():
    from multigrid_jl_b200 import device
    L = device.lib()
    it = ctypes.c_int(0)
    res = np.zeros(4)
    st = L.mgb200_solveMG(None, None, None, ctypes.c_double(1e-6), 3, ctypes.byref(it), res.ctypes.data_as(ctypes.c_void_p))
    assert st == -1 and b"null" in L.mgb200_last_error()


def test_product_package_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "multigrid.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f"{f} imports the oracle"
                assert "libmg_oracle" not in txt and "oracle/" not in txt, f"{f} links the oracle"


@pytest.mark.parametrize("n", [1, 2, 3, 5])
@pytest.mark.parametrize("cplx", [False, True])
def test_host_pinv_matches_numpy(n, cplx):
    """hermitian_pinv_apply (smalldense.h) = Julia/numpy pinv rule on FGMRES_relaxation's H."""
    from multigrid_jl_b200 import device
    L = device.lib()
    rng = np.random.default_rng(n)
    B = rng.standard_normal((n, n)) + (1j * rng.standard_normal((n, n)) if cplx else 0)
    H = B @ B.conj().T
    if n >= 3:                      # rank-deficient case: unused basis columns are zero rows/cols
        H[-1, :] = 0
        H[:, -1] = 0
    xi = rng.standard_normal(n) + (1j * rng.standard_normal(n) if cplx else 0)
    Hc = np.ascontiguousarray(H, dtype=np.complex128)
    xc = np.ascontiguousarray(xi, dtype=np.complex128)
    t = np.zeros(n, dtype=np.complex128)
    assert L.mgb200_host_pinv_apply(n, Hc.ctypes.data_as(ctypes.c_void_p), xc.ctypes.data_as(ctypes.c_void_p),
                                    t.ctypes.data_as(ctypes.c_void_p)) == 0
    ref = np.linalg.pinv(Hc, rcond=np.finfo(float).eps * n) @ xc
    np.testing.assert_allclose(t, ref, rtol=1e-9, atol=1e-12 * np.abs(ref).max())


@pytest.mark.parametrize("cols", [1, 2, 5, 10])
def test_host_hessenberg_lsq_matches_numpy(cols):
    from multigrid_jl_b200 import device
    L = device.lib()
    rng = np.random.default_rng(cols)
    H = np.triu(rng.standard_normal((cols + 1, cols)) + 1j * rng.standard_normal((cols + 1, cols)), -1)
    xi = np.zeros(cols + 1, dtype=np.complex128)
    xi[0] = 2.5
    Hc = np.ascontiguousarray(H)
    y = np.zeros(cols, dtype=np.complex128)
    res = ctypes.c_double(0)
    assert L.mgb200_host_hessenberg_lsq(cols, Hc.ctypes.data_as(ctypes.c_void_p), xi.ctypes.data_as(ctypes.c_void_p),
                                        y.ctypes.data_as(ctypes.c_void_p), ctypes.byref(res)) == 0
    yr = np.linalg.lstsq(Hc, xi, rcond=None)[0]
    np.testing.assert_allclose(y, yr, rtol=1e-10, atol=1e-13)
    assert abs(res.value - np.linalg.norm(Hc @ yr - xi)) < 1e-12


@pytest.mark.parametrize("n,cplx,rank_def", [(1, False, False), (4, False, False), (6, True, False), (5, True, True),
                                             (32, False, False)])
def test_host_general_pinv_matches_numpy(n, cplx, rank_def):
    """general_pinv (one-sided Jacobi SVD) = numpy pinv with the same cut-off; used by blockCG."""
    from multigrid_jl_b200 import device
    L = device.lib()
    rng = np.random.default_rng(n)
    A = rng.standard_normal((n, n)) + (1j * rng.standard_normal((n, n)) if cplx else 0)
    if rank_def:
        A[:, -1] = A[:, 0] * 2.0
    Ac = np.ascontiguousarray(A, dtype=np.complex128)
    P = np.zeros((n, n), dtype=np.complex128)
    assert L.mgb200_host_general_pinv(n, Ac.ctypes.data_as(ctypes.c_void_p), ctypes.c_double(1e-10),
                                      P.ctypes.data_as(ctypes.c_void_p)) == 0
    ref = np.linalg.pinv(Ac, rcond=1e-10)
    np.testing.assert_allclose(P, ref, rtol=1e-8, atol=1e-10 * np.abs(ref).max())


def _vp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


@pytest.mark.parametrize("rows,cols,nb", [(2, 1, 1), (8, 4, 4), (15, 10, 5), (96, 64, 32)])
def test_host_dense_lsq_matches_numpy(rows, cols, nb):
    """dense_lsq (smalldense.h): the block-Hessenberg least-squares problem of blockFGMRES."""
    from multigrid_jl_b200 import device
    L = device.lib()
    rng = np.random.default_rng(rows)
    A = np.ascontiguousarray(rng.standard_normal((rows, cols)) + 1j * rng.standard_normal((rows, cols)))
    B = np.ascontiguousarray(rng.standard_normal((rows, nb)) + 1j * rng.standard_normal((rows, nb)))
    Y = np.zeros((cols, nb), dtype=np.complex128)
    res = ctypes.c_double(0)
    assert L.mgb200_host_dense_lsq(rows, cols, nb, _vp(A), _vp(B), _vp(Y), ctypes.byref(res)) == 0
    Yr = np.linalg.lstsq(A, B, rcond=None)[0]
    np.testing.assert_allclose(Y, Yr, rtol=1e-9, atol=1e-12)
    assert abs(res.value - np.linalg.norm(A @ Yr - B)) < 1e-10


@pytest.mark.parametrize("m,cplx", [(1, False), (4, False), (7, True), (32, True)])
def test_host_cholesky_and_lu_solve(m, cplx):
    """cholesky_upper / lu_solve_small (smalldense.h): Cholesky QR of blockFGMRES, m x m systems of blockBiCGSTB."""
    from multigrid_jl_b200 import device
    L = device.lib()
    rng = np.random.default_rng(m)
    C = rng.standard_normal((m + 3, m)) + (1j * rng.standard_normal((m + 3, m)) if cplx else 0)
    G = np.ascontiguousarray(C.conj().T @ C, dtype=np.complex128)
    R = np.zeros((m, m), dtype=np.complex128)
    assert L.mgb200_host_small_factor(0, m, 0, _vp(G), None, _vp(R)) == 0
    assert np.allclose(np.tril(R, -1), 0) and np.all(np.diag(R).real > 0) and np.allclose(np.diag(R).imag, 0)
    np.testing.assert_allclose(R.conj().T @ R, G, rtol=1e-12, atol=1e-12 * np.abs(G).max())
    G[0, 0] = -1.0
    assert L.mgb200_host_small_factor(0, m, 0, _vp(G), None, _vp(R)) == -5       # not positive definite
    A = np.ascontiguousarray(rng.standard_normal((m, m)) + (1j * rng.standard_normal((m, m)) if cplx else 0),
                             dtype=np.complex128)
    B = np.ascontiguousarray(rng.standard_normal((m, 3)) + 0j)
    X = np.zeros((m, 3), dtype=np.complex128)
    assert L.mgb200_host_small_factor(1, m, 3, _vp(A), _vp(B), _vp(X)) == 0
    np.testing.assert_allclose(A @ X, B, rtol=1e-9, atol=1e-10)
    if m > 1:
        A[:, -1] = A[:, 0]
        assert L.mgb200_host_small_factor(1, m, 3, _vp(A), _vp(B), _vp(X)) == -5   # singular


def test_host_interior_rows_of_a_slab():
    """interior_rows (dist.cuh): rows of a z-slab that read no ghost plane - everything but the first and last
    plane for a middle slab, nothing cut at a domain end, no interior when a middle row reads both sides."""
    import scipy.sparse as sp
    import multigrid_jl_b200 as mg
    from multigrid_jl_b200 import device
    L = device.lib()
    M = mg.getRegularMesh([0, 1, 0, 1, 0, 1], [4, 4, 11])
    A = sp.csr_matrix(mg.poisson_shifted(M, 1e-4))
    plane = 25
    out = np.zeros(2, dtype=np.int64)
    for lo_pl, hi_pl, expect in ((0, 4, (0, 3 * plane)), (4, 8, (plane, 3 * plane)), (8, 12, (plane, 4 * plane))):
        S = A[lo_pl * plane:hi_pl * plane]
        cols = np.ascontiguousarray(S.indices.astype(np.int64) - lo_pl * plane)
        ptr = np.ascontiguousarray(S.indptr.astype(np.int64))
        n_owned = (hi_pl - lo_pl) * plane
        assert L.mgb200_host_interior_rows(ctypes.c_int64(S.shape[0]), _vp(ptr), _vp(cols), ctypes.c_int64(n_owned),
                                           _vp(out)) == 0
        assert tuple(out) == expect
    ptr = np.array([0, 1, 3, 4], dtype=np.int64)
    cols = np.array([0, -1, 5, 1], dtype=np.int64)     # row 1 reads a lower and an upper ghost
    assert L.mgb200_host_interior_rows(ctypes.c_int64(3), _vp(ptr), _vp(cols), ctypes.c_int64(3), _vp(out)) == 0
    assert out[0] == out[1]
