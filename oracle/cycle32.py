"""TEST INFRASTRUCTURE - CPU restatement of the reference's cycle for SINGLE-PRECISION hierarchies
(getMGparam(Float32 / ComplexF32, ...): `singlePrecision`, MGdef.jl:119,151; MGsetup.jl:31-33,79-82,108-110) and of
the mixed-precision preconditioner closure (getMultigridPreconditioner, SolveFuncs.jl:52-60).

Only tests/ may import this.  Same operation order as oracle/cycle.py (which restates MGcycle.jl:1-136 for the
double-precision types through the C kernels); here every array lives in the hierarchy's value type and the passes
are plain scipy / numpy operations in that type: one rounding per product and per sum, rows accumulated in stored
order.  Diagonal smoothers ("Jac", "SPAI") and V / F / W cycles.

PARITY UNPINNED, like the rest of oracle/ (no runnable reference, no golden vectors in its tests).  One point is a
recollection of Julia 1.7 behaviour and not of the reference's source: `lu` of a Float32 / ComplexF32 sparse matrix
promotes it to double precision (UMFPACK has no single-precision factorisation), `LU \\ b` is then computed in double
precision and `x[:] = z` rounds it (MGcycle.jl:176-179).
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp


def _wide(dt):
    return np.complex128 if np.dtype(dt).kind == "c" else np.float64


class OracleMG32:
    def __init__(self, param):
        # written for the single-precision types; it is type-generic, and tests/test_oracle.py also runs it in double
        # precision as an independent (scipy / numpy) cross-check of oracle/cycle.py + mg_kernels.c
        self.VAL = np.dtype(param.VAL)
        assert self.VAL in (np.dtype(np.float32), np.dtype(np.complex64), np.dtype(np.float64), np.dtype(np.complex128))
        self.relaxPre, self.relaxPost = param.relaxPre, param.relaxPost
        self.cycleType = param.cycleType
        self.maxOuterIter, self.relativeTol = param.maxOuterIter, param.relativeTol
        assert param.relaxType in ("Jac", "SPAI") and param.coarseSolveType not in ("GMRES",)
        # operators in CSR (the stored CSC arrays of the adjoint are the CSR arrays of conj(operator))
        self.A = [sp.csr_matrix(sp.csc_matrix(M).conj().T, dtype=self.VAL) for M in param.As]
        rT = np.zeros(0, dtype=self.VAL).real.dtype
        self.P = [sp.csr_matrix(sp.csc_matrix(M).T, dtype=rT) for M in param.Ps]
        self.R = [sp.csr_matrix(sp.csc_matrix(M).T, dtype=rT) for M in param.Rs]
        for M in self.A + self.P + self.R:
            M.sort_indices()
        self.d = [np.ascontiguousarray(d, dtype=self.VAL) for d in param.relaxPrecs]
        self.LU = sla.lu_factor(self.A[-1].toarray().astype(_wide(self.VAL)))

    # scipy multiplies a real single-precision matrix with a complex single-precision vector in complex64
    def mul(self, M, x):
        y = M @ x
        assert y.dtype == self.VAL, (y.dtype, self.VAL)
        return y


def relax(mg, A, r, x, b, d, numit):
    """MGcycle.jl:122-136 (r is stale on return)."""
    dd = d if x.ndim == 1 else d[:, None]
    for _ in range(numit - 1):
        x += dd * r
        r = -mg.mul(A, x)
        r += b
    x += dd * r
    return x


def recursiveCycle(mg: OracleMG32, b, x, level):
    """MGcycle.jl:1-118 (1-based level), diagonal smoothers, V / F / W."""
    L = len(mg.A)
    if level == L:
        z = sla.lu_solve(mg.LU, b.astype(_wide(mg.VAL)))
        x[...] = z.astype(mg.VAL)
        return x
    A, d = mg.A[level - 1], mg.d[level - 1]
    r = b.copy()
    if np.linalg.norm(x) > 0.0:
        r -= mg.mul(A, x)
    x = relax(mg, A, r, x, b, d, mg.relaxPre(level))
    r = -mg.mul(A, x)
    r += b
    bc = mg.mul(mg.R[level - 1], r)
    xc = np.zeros_like(bc)
    if level == L - 1:
        xc = recursiveCycle(mg, bc, xc, level + 1)
    else:
        xc = recursiveCycle(mg, bc, xc, level + 1)
        if mg.cycleType == 'W':
            xc = recursiveCycle(mg, bc, xc, level + 1)
        elif mg.cycleType == 'F':
            mg.cycleType = 'V'
            xc = recursiveCycle(mg, bc, xc, level + 1)
            mg.cycleType = 'F'
    x += mg.mul(mg.P[level - 1], xc)
    r = b.copy()
    r -= mg.mul(A, x)
    x = relax(mg, A, r, x, b, d, mg.relaxPost(level))
    return x


def solveMG(mg: OracleMG32, b, x):
    """SolveFuncs.jl:3-39 with single-precision b, x: (x, iter, resvec)."""
    b = np.asarray(b, dtype=mg.VAL)
    x = np.array(x, dtype=mg.VAL)
    A = mg.A[0]
    if np.linalg.norm(x) == 0:
        res = float(np.linalg.norm(b.astype(_wide(mg.VAL))))
    else:
        res = float(np.linalg.norm((b - mg.mul(A, x)).astype(_wide(mg.VAL))))
    res_init = res
    resvec = [res_init]
    it = 0
    for _ in range(mg.maxOuterIter):
        x = recursiveCycle(mg, b, x, 1)
        r = -mg.mul(A, x)
        r += b
        it += 1
        res = float(np.linalg.norm(r.astype(_wide(mg.VAL))))
        resvec.append(res)
        if res / res_init < mg.relativeTol:
            break
    return x, it, np.asarray(resvec)


def getMultigridPreconditioner(mg: OracleMG32):
    """SolveFuncs.jl:52-60, mixed_precision branch: bl[:] .= b (rounded); z .= 0; recursiveCycle; z2[:] .= z."""
    def MMG(b):
        bl = np.asarray(b).astype(mg.VAL)
        z = np.zeros_like(bl)
        z = recursiveCycle(mg, bl, z, 1)
        return z.astype(_wide(mg.VAL))
    return MMG
