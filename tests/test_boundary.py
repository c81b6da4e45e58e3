"""The drop-in boundary as the reference would use it: 1-based Int64 colptr / rowval exactly as a
Julia SparseMatrixCSC holds them (parRelax.jl:61-79), a C++ caller compiled against include/mgb200.h alone, and the
per-call Krylov matrix of solveCG_MG(AT, ...) (SolveFuncs.jl:73-82)."""
import os
import shutil
import subprocess

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import make_problem

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLIENT_SRC = os.path.join(ROOT, "tests", "abi_client", "abi_client.cpp")


def _build_client(tmp_path):
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    if gxx is None:
        pytest.skip("no C++ compiler")
    exe = str(tmp_path / "abi_client")
    pkg = os.path.join(ROOT, "multigrid.jl_b200")
    subprocess.check_call([gxx, "-std=c++14", "-O1", "-I", os.path.join(ROOT, "include"), CLIENT_SRC, "-o", exe,
                           "-L", pkg, "-l:libmgb200.so", "-Wl,-rpath," + pkg])
    return exe


def test_cpp_client_links_against_the_header(tmp_path):
    """A C++ program that includes only include/mgb200.h compiles, links and loads; without a GPU mgb200_create
    reports an error status (no CPU fallback) instead of crashing."""
    exe = _build_client(tmp_path)
    out = subprocess.run([exe, "probe"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    assert "version" in out.stdout and "null-handle status -1" in out.stdout
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        assert "create status -2" in out.stdout


def _write_csc(f, M, dtype=np.float64):
    M = sp.csc_matrix(M)
    M.sort_indices()
    (M.indptr.astype(np.int64) + 1).tofile(f)        # Julia: 1-based
    (M.indices.astype(np.int64) + 1).tofile(f)
    np.ascontiguousarray(M.data, dtype=dtype).tofile(f)


@pytest.mark.gpu
def test_cpp_client_solves_with_julia_arrays(tmp_path):
    """The C++ caller feeds 1-based Int64 arrays through mgb200_upload_level(..., index_base = 1) and gets the
    residual histories and solutions of the ctypes path (0-based) bit for bit."""
    import multigrid_jl_b200 as mg
    exe = _build_client(tmp_path)
    A, AT, M, p, b = make_problem("poisson", [24, 20, 16], 3, maxit=8, tol=1e-9)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        np.array([len(p.As), ord('V'), p.maxOuterIter], dtype=np.int64).tofile(f)
        np.array([p.relativeTol]).tofile(f)
        for l in range(len(p.As) - 1):
            n, nc = p.As[l].shape[1], p.As[l + 1].shape[1]
            np.array([n, nc, p.As[l].nnz, p.Ps[l].nnz, p.Rs[l].nnz], dtype=np.int64).tofile(f)
            _write_csc(f, p.As[l]); _write_csc(f, p.Ps[l]); _write_csc(f, p.Rs[l])
            np.ascontiguousarray(p.relaxPrecs[l], dtype=np.float64).tofile(f)
        np.array([p.As[-1].shape[1], p.As[-1].nnz], dtype=np.int64).tofile(f)
        _write_csc(f, p.As[-1])
        b.tofile(f)
    out = subprocess.run([exe, "solve", fin, fout], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    raw = np.fromfile(fout, dtype=np.uint8)
    head = raw[:24].view(np.int64)
    k = p.maxOuterIter + 1
    body = raw[24:].view(np.float64)
    res, rescg, x, xcg = body[:k], body[k:2 * k], body[2 * k:2 * k + len(b)], body[2 * k + len(b):]
    # the same through ctypes with 0-based arrays
    x0 = np.zeros_like(b)
    _, _, it = mg.solveMG(p, b, x0)
    res0 = np.array(p.last_resvec)
    xc0 = np.zeros_like(b)
    _, _, itcg = mg.solveCG_MG(AT, p, b, xc0)
    rescg0 = np.array(p.last_resvec)
    assert head[0] == it and head[1] == itcg and head[2] == p.last_flag
    assert np.array_equal(res[:it + 1], res0[:it + 1]) and np.array_equal(x, x0)
    assert np.array_equal(rescg[:itcg], rescg0[:itcg]) and np.array_equal(xcg, xc0)


@pytest.mark.gpu
@pytest.mark.parametrize("kind,n", [("poisson", [24, 20, 16]), ("helmholtz", [40, 36]), ("diffusion", [20, 20, 12])])
def test_one_based_upload_bit_identical(kind, n):
    """index_base = 1 (what AT.colptr / AT.rowval are in Julia) against index_base = 0: identical device hierarchies."""
    import multigrid_jl_b200 as mg
    A, AT, M, p, b = make_problem(kind, n, 3, maxit=6)
    out = []
    for base in (0, 1):
        dev = mg.DeviceHierarchy(p, device=0, index_base=base)
        dev.set_cycle(p)
        if base == 1:
            dev.set_krylov_matrix(AT, index_base=1)
        x, it, res = dev.solveMG(b, np.zeros_like(b), p.relativeTol, p.maxOuterIter)
        xk, itk, flag, resk = dev.solveFGMRES(b, np.zeros_like(b), 5, True, 1e-8, 10)
        out.append((x.copy(), it, res.copy(), xk.copy(), itk, flag, resk.copy()))
        dev.destroy()
    a, c = out
    assert a[1] == c[1] and a[4] == c[4] and a[5] == c[5]
    assert np.array_equal(a[2], c[2]) and np.array_equal(a[0], c[0])
    assert np.array_equal(a[6], c[6]) and np.array_equal(a[3], c[3])


@pytest.mark.gpu
def test_krylov_matrix_is_per_call():
    """solveCG_MG(AT2, param, ...) followed by solveCG_MG(As[1], param, ...) on the same param: the second solve must
    multiply with As[1] again (the reference builds Afun from the AT of every call, SolveFuncs.jl:73-82)."""
    import multigrid_jl_b200 as mg
    A, AT, M, p, b = make_problem("poisson", [32, 32], 3, maxit=40, tol=1e-8)
    x1 = np.zeros_like(b)
    mg.solveCG_MG(AT, p, b, x1)
    res1, it1 = np.array(p.last_resvec), len(p.last_resvec)
    AT2 = (AT + 0.05 * abs(AT).max() * sp.identity(AT.shape[0], format="csc")).tocsc()
    x2 = np.zeros_like(b)
    mg.solveCG_MG(AT2, p, b, x2)
    assert np.linalg.norm(b - AT2.T @ x2) <= 1.1e-8 * np.linalg.norm(b)
    assert np.linalg.norm(b - AT.T @ x2) > 1e-4 * np.linalg.norm(b)      # it really solved the other system
    x3 = np.zeros_like(b)
    mg.solveCG_MG(AT, p, b, x3)                    # back to the hierarchy's own matrix
    assert len(p.last_resvec) == it1 and np.array_equal(np.array(p.last_resvec), res1) and np.array_equal(x3, x1)
    x4 = np.zeros_like(b)
    mg.solveGMRES_MG(p.As[0], p, b, x4, True, 5)
    assert np.linalg.norm(b - AT.T @ x4) <= 1.1e-8 * np.linalg.norm(b)
