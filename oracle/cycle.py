"""TEST INFRASTRUCTURE - CPU restatement of the reference's multigrid cycle.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this.  Every function restates the cited lines of
JuliaInv/Multigrid.jl v0.8.0 in the reference's own operation order and WITHOUT
fusion (separate SpMatMul, addVectors, scale-add and norm passes); the array
loops run in oracle/mg_kernels.c (OpenMP over rows, as ParSpMatVec does).

PARITY UNPINNED: the reference cannot be executed here (no julia/gfortran) and
its tests hold no golden vectors, known-answer tests or iteration-count
assertions for this path.  What pins this oracle is
(i) the cited source lines, (ii) scipy cross-checks of each kernel in
tests/test_oracle.py and (iii) the reference tests' residual thresholds
re-run on seeded inputs (tests/test_reference_thresholds.py).
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp

from . import kernels as K


class CYCLEmem:
    """MGdef.jl:56-60"""

    def __init__(self, n, m, T, withB=True):
        shape = (n,) if m == 1 else (n, m)
        self.b = np.zeros(shape if withB else (0,), dtype=T, order="F")
        self.r = np.zeros(shape, dtype=T, order="F")
        self.x = np.zeros(shape, dtype=T, order="F")


class FGMRESmem:
    """FGMRES.jl:3-29"""

    def __init__(self, n, T, k, m=1):
        shape = (n,) if m == 1 else (n, m)
        self.v_prec = np.zeros(shape, dtype=T, order="F")
        self.Az = np.zeros(shape, dtype=T, order="F")
        self.V = np.zeros((n * m, k), dtype=T, order="F")
        self.Z = np.zeros((n * m, k), dtype=T, order="F")

    def reset(self):
        self.v_prec[...] = 0
        self.Az[...] = 0
        self.V[...] = 0
        self.Z[...] = 0


class OracleMG:
    """Host copy of an MGparam hierarchy in the reference's storage convention
    plus the workspaces of adjustMemoryForNumRHS (MGsetup.jl:166-223)."""

    def __init__(self, param, numCores=None):
        self.VAL = np.dtype(param.VAL)
        self.levels = param.levels
        self.numCores = K.max_threads() if numCores is None else int(numCores)
        self.maxOuterIter = param.maxOuterIter
        self.relativeTol = param.relativeTol
        self.relaxType = param.relaxType
        self.relaxParam = param.relaxParam
        self.relaxPre = param.relaxPre
        self.relaxPost = param.relaxPost
        self.cycleType = param.cycleType
        self.coarseSolveType = param.coarseSolveType
        self.As = [K.CSCAdjoint(A) for A in param.As]
        self.Ps = [K.CSCAdjoint(P) for P in param.Ps]
        self.Rs = [K.CSCAdjoint(R) for R in param.Rs]
        self.relaxPrecs = [np.ascontiguousarray(d, dtype=self.VAL) for d in param.relaxPrecs]
        self.memCycle = []
        self.memRelax = []
        self.memKcycle = []
        self.counters = {"A": 0, "P": 0, "R": 0, "coarse": 0}
        # defineCoarsestAinv (MGsetup.jl:350): LU of sparse(AT') = A_L.  UMFPACK is replaced by a
        # dense LAPACK LU with partial pivoting (the device uses the same class of factorisation).
        if self.coarseSolveType == "GMRES":
            self.LU = np.ascontiguousarray(param.LU, dtype=self.VAL)
        else:
            AL = sp.csc_matrix(param.As[-1]).conj().T.toarray().astype(self.VAL)
            self.LU = sla.lu_factor(AL)

    # -- MGsetup.jl:166-223 ------------------------------------------------------------------
    def adjustMemoryForNumRHS(self, nrhs=1):
        if self.memCycle and (1 if self.memCycle[0].x.ndim == 1 else self.memCycle[0].x.shape[1]) == nrhs:
            return self
        T = self.VAL
        L = self.levels
        self.memKcycle = [None] * max(L - 2, 0) if self.cycleType == 'K' else []
        self.memRelax = [None] * (L - 1) if self.relaxType == "Jac-GMRES" else []
        self.memCycle = [None] * L
        for l in range(1, L):  # 1-based level l
            N = self.As[l - 1].shape[1]
            self.memCycle[l - 1] = CYCLEmem(N, nrhs, T, True)
            if self.relaxType == "Jac-GMRES":
                maxRelax = max(self.relaxPre(l), self.relaxPost(l))
                self.memRelax[l - 1] = FGMRESmem(N, T, maxRelax, nrhs)
            if l > 1 and self.cycleType == 'K':
                self.memKcycle[l - 2] = FGMRESmem(N, T, 2, nrhs)
        self.memCycle[-1] = CYCLEmem(self.As[-1].shape[1], nrhs, T, False)
        self.memCycle[-1].b = self.memCycle[-1].r  # aliased (MGsetup.jl:217-218)
        return self


def getAfun(AT: K.CSCAdjoint, Az, numCores, counter=None):
    """SolveFuncs.jl:65-71: z -> A z, returning the same (aliased) buffer."""
    def Afun(z):
        if counter is not None:
            counter["A"] += 1
        K.SpMatMul4(AT, z, Az, numCores)
        return Az
    return Afun


def relax(mg, AT, r, x, b, d, numit, numCores):
    """MGcycle.jl:122-136.  r is stale on return."""
    for _ in range(numit - 1):
        K.scaleAdd(d, r, x, numCores)                 # x .+= d.*r
        K.SpMatMul(-1.0, AT, x, 0.0, r, numCores)     # r = -A'*x
        mg.counters["A"] += 1
        K.addVectors(1.0, b, r, numCores)             # r = r + b
    K.scaleAdd(d, r, x, numCores)
    return x


def solveCoarsest(mg, b, x):
    """MGcycle.jl:138-181: default branch ``z = param.LU\\b; x[:] = z`` and the
    "GMRES" branch (one restart of fgmres(10), tol 0.01, Jacobi preconditioner)."""
    mg.counters["coarse"] += 1
    if mg.coarseSolveType == "GMRES":
        from . import krylov
        AT = mg.As[-1]
        Afun = getAfun(AT, np.zeros(b.shape, dtype=b.dtype, order="F"), mg.numCores)
        d = mg.LU
        y = np.zeros(b.shape, dtype=b.dtype, order="F")

        def M2(xx):
            y[...] = (d * xx) if xx.ndim == 1 else (d[:, None] * xx)
            return y
        x[...] = 0.0
        if b.ndim == 1:
            xs = krylov.fgmres(Afun, b, 10, tol=0.01, maxIter=1, M=M2, x=x)[0]
        else:
            xs = krylov.blockFGMRES(Afun, b, 10, tol=0.01, maxIter=1, M=M2, X=x)[0]
        x[...] = xs
        return x
    z = sla.lu_solve(mg.LU, b)
    x[...] = z
    return x


def FGMRES_relaxation(Afun, r0, x0, inner, prec, TOL, numCores, mem=None):
    """FGMRES.jl:48-126 (multi-RHS blocks are flattened column-major into one vector)."""
    T = r0.dtype
    m = 1 if r0.ndim == 1 else r0.shape[1]
    n = r0.shape[0]
    if mem is None:
        mem = FGMRESmem(n, T, inner, m)
    else:
        mem.reset()
        if mem.V.shape[1] != inner:
            raise RuntimeError("FGMRES_relaxation: size of Krylov subspace is different than inner")
    rnorm0 = K.norm(r0, numCores)
    H = np.zeros((inner, inner), dtype=T)
    xi = np.zeros(inner, dtype=T)
    t = np.zeros(inner, dtype=T)
    Z, AZ = mem.Z, mem.V
    w = None
    rnorms = np.zeros(inner)
    nj = inner
    for j in range(inner):
        z = prec(r0) if j == 0 else prec(w)
        Z[:, j] = z.reshape(n * m, order="F")
        w = Afun(z)
        wf = w.reshape(n * m, order="F")
        AZ[:, j] = wf
        for i in range(inner):                        # BLAS.gemv!('C',1,AZ,w,0,t)
            t[i] = K.dot(np.ascontiguousarray(AZ[:, i]), wf, numCores)
        xi[j] = K.dot(wf, r0.reshape(n * m, order="F"), numCores)
        H[:, j] = t
        H[j, :] = np.conj(t)
        H = 0.5 * H + 0.5 * H.conj().T
        t[:] = julia_pinv(H) @ xi
        rnorms[j] = np.sqrt(abs((np.vdot(t, H @ t) - 2.0 * np.vdot(t, xi) + rnorm0 ** 2).real))
        if rnorms[j] < TOL:
            nj = j + 1
            break
    rnorms = rnorms[:nj]
    wcorr = (Z @ t).reshape(x0.shape, order="F")      # BLAS.gemv!('N',1,Z,t,0,w)
    wcorr = np.asfortranarray(wcorr)
    K.addVectors(1.0, wcorr, x0, numCores)
    return x0, rnorms


def julia_pinv(H):
    """LinearAlgebra.pinv default: SVD, singular values <= eps*min(size)*max(S) are dropped
    (Julia 1.7 stdlib; recollection)."""
    H = np.atleast_2d(H)
    rtol = np.finfo(np.float64).eps * min(H.shape)
    return np.linalg.pinv(H, rcond=rtol)


def recursiveCycle(mg: OracleMG, b, x, level):
    """MGcycle.jl:1-118; ``level`` is 1-based as in the reference."""
    gmresTol = 1e-5
    numCores = mg.numCores
    nlevels = len(mg.As)
    if level == nlevels:
        r = mg.memCycle[level - 1].r
        r[...] = b
        return solveCoarsest(mg, r, x)
    AT = mg.As[level - 1]
    r = mg.memCycle[level - 1].r
    r[...] = b
    if K.norm(x, numCores) > 0.0:
        K.SpMatMul(-1.0, AT, x, 1.0, r, numCores)     # r -= A'*x
        mg.counters["A"] += 1
    D = mg.relaxPrecs[level - 1]
    PT = mg.Ps[level - 1]
    RT = mg.Rs[level - 1]
    npresmth = mg.relaxPre(level)
    npostsmth = mg.relaxPost(level)
    if mg.relaxType == "Jac-GMRES":
        memR = mg.memRelax[level - 1]
        y = memR.v_prec

        def MM(xx):
            y[...] = (D * xx) if xx.ndim == 1 else (D[:, None] * xx)
            return y
        Afun = getAfun(AT, memR.Az, numCores, mg.counters)
        x = FGMRES_relaxation(Afun, r, x, npresmth, MM, gmresTol, numCores, memR)[0]
    else:
        x = relax(mg, AT, r, x, b, D, npresmth, numCores)
    K.SpMatMul(-1.0, AT, x, 0.0, r, numCores)         # r = -A'*x
    mg.counters["A"] += 1
    K.addVectors(1.0, b, r, numCores)                 # r = r + b
    xc = mg.memCycle[level].x
    xc[...] = 0.0
    bc = mg.memCycle[level].b
    bc = K.SpMatMul4(RT, r, bc, numCores)
    mg.counters["R"] += 1
    if level == nlevels - 1:
        xc = solveCoarsest(mg, bc, xc)
    else:
        Ac = mg.As[level]
        if mg.cycleType == 'K':
            memK = mg.memKcycle[level - 1]
            yzK = memK.v_prec
            AfunK = getAfun(Ac, memK.Az, numCores, mg.counters)

            def MMG(v):
                yzK[...] = 0.0
                return recursiveCycle(mg, v, yzK, level + 1)
            xc = FGMRES_relaxation(AfunK, bc, xc, 2, MMG, gmresTol, numCores, memK)[0]
        else:
            xc = recursiveCycle(mg, bc, xc, level + 1)
            if mg.cycleType == 'W':
                xc = recursiveCycle(mg, bc, xc, level + 1)
            elif mg.cycleType == 'F':
                mg.cycleType = 'V'
                xc = recursiveCycle(mg, bc, xc, level + 1)
                mg.cycleType = 'F'
    K.SpMatMul(1.0, PT, xc, 1.0, x, numCores)         # x += PT'*xc
    mg.counters["P"] += 1
    r[...] = b
    K.SpMatMul(-1.0, AT, x, 1.0, r, numCores)         # r -= A'*x
    mg.counters["A"] += 1
    if mg.relaxType == "Jac-GMRES":
        Afun = getAfun(AT, mg.memRelax[level - 1].Az, numCores, mg.counters)
        x = FGMRES_relaxation(Afun, r, x, npostsmth, MM, gmresTol, numCores, mg.memRelax[level - 1])[0]
    else:
        x = relax(mg, AT, r, x, b, D, npostsmth, numCores)
    return x


def solveMG(mg: OracleMG, b, x, verbose=False):
    """SolveFuncs.jl:3-39.  Returns (x, iter, resvec) where resvec[0] = res_init and
    resvec[k] = ||b - A x_k|| after cycle k (the pinned per-cycle quantity)."""
    b = np.asfortranarray(b)
    nrhs = 1 if b.ndim == 1 else b.shape[1]
    mg.adjustMemoryForNumRHS(nrhs)
    tol = mg.relativeTol
    numCores = mg.numCores
    AT = mg.As[0]
    r = mg.memCycle[0].r
    r[...] = b
    if K.norm(x, numCores) == 0:
        res = K.norm(b, numCores)
    else:
        K.SpMatMul(-1.0, AT, x, 1.0, r, numCores)
        res = K.norm(r, numCores)
    res_init = res
    resvec = [res_init]
    it = 0
    for count in range(1, mg.maxOuterIter + 1):
        x = recursiveCycle(mg, b, x, 1)
        K.SpMatMul(-1.0, AT, x, 0.0, r, numCores)
        K.addVectors(1.0, b, r, numCores)
        it += 1
        res_prev = res
        res = K.norm(r, numCores)
        resvec.append(res)
        if verbose:
            print(f"Cycle {count} done with relres: {res / res_init}. Convergence factor: {res / res_prev}")
        if res / res_init < tol:
            break
    return x, it, np.asarray(resvec)


def getMultigridPreconditioner(mg: OracleMG, B):
    """SolveFuncs.jl:43-63 (same-precision branch): r -> (z .= 0; cycle; z), z aliased."""
    nrhs = 1 if B.ndim == 1 else B.shape[1]
    mg.adjustMemoryForNumRHS(nrhs)
    z = mg.memCycle[0].x

    def MMG(bb):
        z[...] = 0.0
        recursiveCycle(mg, bb, z, 1)
        return z
    return MMG


def solveCG_MG(AT, mg: OracleMG, b, x0):
    """SolveFuncs.jl:77-79,103-116 -> KrylovMethods.cg / blockCG."""
    from . import krylov
    b = np.asfortranarray(b)
    ATc = AT if isinstance(AT, K.CSCAdjoint) else K.CSCAdjoint(AT)
    Afun = getAfun(ATc, np.zeros(b.shape, dtype=b.dtype, order="F"), mg.numCores)
    MMG = getMultigridPreconditioner(mg, b)
    if b.ndim == 1 or b.shape[1] == 1:
        x, flag, rnorm, it, resvec = krylov.cg(Afun, b.reshape(-1), tol=mg.relativeTol,
                                               maxIter=mg.maxOuterIter, M=MMG, x=x0)
    else:
        x, flag, rnorm, it, resvec = krylov.blockCG(Afun, b, tol=mg.relativeTol,
                                                    maxIter=mg.maxOuterIter, M=MMG, X=x0)
    return x, it, flag, resvec


def solveBiCGSTAB_MG(AT, mg: OracleMG, b, x0):
    """SolveFuncs.jl:73-75,85-99 -> KrylovMethods.bicgstb / blockBiCGSTB (M1 = one cycle, M2 = identity).
    Returns (x, iter, flag, resvec, nprec) with nprec = 2*iter*nrhs + (flag == -3)*nrhs (SolveFuncs.jl:97)."""
    from . import krylov
    b = np.asfortranarray(b)
    ATc = AT if isinstance(AT, K.CSCAdjoint) else K.CSCAdjoint(AT)
    Afun = getAfun(ATc, np.zeros(b.shape, dtype=b.dtype, order="F"), mg.numCores)
    MMG = getMultigridPreconditioner(mg, b)
    if b.ndim == 1 or b.shape[1] == 1:
        x, flag, rnorm, it, resvec = krylov.bicgstb(Afun, b.reshape(-1), tol=mg.relativeTol,
                                                    maxIter=mg.maxOuterIter, M1=MMG, x=x0)
        nrhs = 1
    else:
        x, flag, rnorm, it, resvec = krylov.blockBiCGSTB(Afun, b, tol=mg.relativeTol,
                                                         maxIter=mg.maxOuterIter, M1=MMG, x=x0)
        nrhs = b.shape[1]
    return x, it, flag, resvec, 2 * it * nrhs + (nrhs if flag == -3 else 0)


def solveGMRES_MG(AT, mg: OracleMG, b, x0, flexible, inner):
    """SolveFuncs.jl:80-82,120-132 -> KrylovMethods.fgmres / blockFGMRES."""
    from . import krylov
    b = np.asfortranarray(b)
    ATc = AT if isinstance(AT, K.CSCAdjoint) else K.CSCAdjoint(AT)
    Afun = getAfun(ATc, np.zeros(b.shape, dtype=b.dtype, order="F"), mg.numCores)
    MMG = getMultigridPreconditioner(mg, b)
    if b.ndim == 1 or b.shape[1] == 1:
        x, flag, rnorm, it, resvec = krylov.fgmres(Afun, b.reshape(-1), inner, tol=mg.relativeTol,
                                                   maxIter=mg.maxOuterIter, M=MMG, x=x0,
                                                   flexible=flexible)
    else:
        x, flag, rnorm, it, resvec = krylov.blockFGMRES(Afun, b, inner, tol=mg.relativeTol,
                                                        maxIter=mg.maxOuterIter, M=MMG, X=x0,
                                                        flexible=flexible)
    return x, it, flag, resvec
