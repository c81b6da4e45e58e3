"""TEST INFRASTRUCTURE - ctypes binding of oracle/libmg_oracle.so (see mg_kernels.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this.  PARITY UNPINNED (no runnable reference, no golden vectors in the
reference's tests): see the header of mg_kernels.c.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False):
    so = os.path.join(_HERE, "libmg_oracle.so")
    src = os.path.join(_HERE, "mg_kernels.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libmg_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.oracle_max_threads.restype = ctypes.c_int64
        _LIB.nrm2_f64.restype = ctypes.c_double
        _LIB.nrm2_c64.restype = ctypes.c_double
    return _LIB


def max_threads() -> int:
    return int(lib().oracle_max_threads())


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _suf(dt):
    return "c64" if np.dtype(dt).kind == "c" else "f64"


class CSCAdjoint:
    """The reference's stored form: a CSC matrix of the ADJOINT operator with
    1-based Int64 index arrays (what Julia hands to ccall)."""

    def __init__(self, M):
        import scipy.sparse as sp
        M = sp.csc_matrix(M)
        if not M.has_sorted_indices:
            M.sort_indices()
        self.shape = M.shape
        self.colptr = (M.indptr.astype(np.int64) + 1)
        self.rowval = (M.indices.astype(np.int64) + 1)
        dt = np.complex128 if np.iscomplexobj(M.data) else np.float64
        self.nzval = np.ascontiguousarray(M.data, dtype=dt)
        self.dtype = np.dtype(dt)

    @property
    def nnz(self):
        return self.nzval.shape[0]


def SpMatMul(alpha, AT: CSCAdjoint, x, beta, target, numCores: int):
    """target = beta*target + alpha*AT'*x  (SpMatMul.jl:4-13); x, target are
    n or n x m Fortran-ordered arrays."""
    assert x.flags.f_contiguous and target.flags.f_contiguous
    assert x.shape[0] == AT.shape[0] and target.shape[0] == AT.shape[1]
    vt = np.dtype(x.dtype)
    assert target.dtype == vt
    nrhs = 1 if x.ndim == 1 else x.shape[1]
    name = f"spmatmul_{_suf(AT.dtype)}_{_suf(vt)}"
    a = np.array([alpha], dtype=vt)
    b = np.array([beta], dtype=vt)
    getattr(lib(), name)(ctypes.c_int64(AT.shape[1]), _p(AT.colptr), _p(AT.rowval), _p(AT.nzval),
                         _p(x), ctypes.c_int64(x.shape[0]), _p(target),
                         ctypes.c_int64(target.shape[0]), ctypes.c_int64(nrhs), _p(a), _p(b),
                         ctypes.c_int64(numCores))
    return target


def SpMatMul4(AT: CSCAdjoint, x, target, numCores: int):
    """4-argument form (SpMatMul.jl:16-26): alpha = 1, beta = 0."""
    return SpMatMul(1.0, AT, x, 0.0, target, numCores)


def addVectors(alpha, x, target, numCores: int = 0):
    """target += alpha*x (SpMatMul.jl:29-36)."""
    assert x.dtype == target.dtype and x.size == target.size
    a = np.array([alpha], dtype=x.dtype)
    getattr(lib(), f"addvectors_{_suf(x.dtype)}")(ctypes.c_int64(x.size), _p(a), _p(x), _p(target),
                                                   ctypes.c_int64(numCores))


def scaleAdd(d, r, x, numCores: int = 0):
    """x .+= d .* r (MGcycle.jl:129,134), d broadcast over the columns."""
    n = d.shape[0]
    nrhs = r.size // n
    getattr(lib(), f"scaleadd_{_suf(x.dtype)}")(ctypes.c_int64(n), ctypes.c_int64(nrhs), _p(d), _p(r),
                                                 _p(x), ctypes.c_int64(numCores))


def norm(x, numCores: int = 0) -> float:
    return float(getattr(lib(), f"nrm2_{_suf(x.dtype)}")(ctypes.c_int64(x.size), _p(x),
                                                          ctypes.c_int64(numCores)))


def dot(x, y, numCores: int = 0):
    """dot(x,y) = sum(conj(x).*y)."""
    out = np.zeros(1, dtype=x.dtype)
    getattr(lib(), f"dot_{_suf(x.dtype)}")(ctypes.c_int64(x.size), _p(x), _p(y), _p(out),
                                            ctypes.c_int64(numCores))
    return out[0]
