"""jInv ``AbstractSolver`` plug-ins of the reference, host side: ``MGsolver`` (src/Multigrid/MGWrapper.jl:6-104)
and ``SA_AMGsolver`` (src/Multigrid/SAAMGWrapper.jl:6-95).  Lazy setup on the first solve, transposed
solves through ``transposeHierarchy``, iteration and time counters; the solve itself runs on the device
through the Krylov drivers of ``solve.py``.  ``solveLinearSystem`` is the reference's
``solveLinearSystem!(A,B,X,param,doTranspose)`` (X is updated in place and returned with param)."""
from __future__ import annotations

import time
import warnings
from dataclasses import dataclass

import numpy as np
import scipy.sparse as sp

from . import mgdef
from .mesh import RegularMesh
from .mgdef import MGparam, hierarchyExists
from .mgsetup import MGsetup, transposeHierarchy
from .sa_amg import SA_AMGsetup
from .solve import solveBiCGSTAB_MG, solveCG_MG, solveGMRES_MG, solveMG


@dataclass
class MGsolver:
    """MGWrapper.jl:6-19."""
    MG: MGparam
    Krylov: str = "GMRES"
    sym: int = 0            # 0 = unsymmetric, 1 = symm. pos. def., 2 = general symmetric
    out: int = -1
    isTranspose: bool = False
    doClear: int = 0
    tol: float = 0.0
    nIter: int = 0
    timeSetup: float = 0.0
    timeSolve: float = 0.0
    Mesh: RegularMesh = None


@dataclass
class SA_AMGsolver:
    """SAAMGWrapper.jl:6-17."""
    MG: MGparam
    Krylov: str = "BiCGSTAB"
    sym: int = 1
    out: int = -1
    isTranspose: bool = False
    doClear: int = 0
    tol: float = 0.0
    nIter: int = 0
    timeSetup: float = 0.0
    timeSolve: float = 0.0


def getMGsolver(MG: MGparam, Mesh: RegularMesh, sym, Krylov: str = "GMRES", out: int = -1) -> MGsolver:
    """MGWrapper.jl:22-25."""
    MG.Meshes = [Mesh]
    return MGsolver(MG, Krylov, int(sym), out, False, 0, MG.relativeTol, 0, 0.0, 0.0, Mesh)


def getSA_AMGsolver(MG: MGparam, Krylov: str = "BiCGSTAB", sym: int = 1, out: int = -1) -> SA_AMGsolver:
    """SAAMGWrapper.jl:20-25."""
    if sym != 1:
        warnings.warn("Non-symmetric AMG version is not implemented yet...")
    return SA_AMGsolver(MG, Krylov, int(sym), out, False, 0, MG.relativeTol, 0, 0.0, 0.0)


def _adjoint(A):
    A = sp.csc_matrix(A)
    return sp.csc_matrix(A.conj().T) if np.iscomplexobj(A.data) else sp.csc_matrix(A.T)


def _prepare(A, B, X, param, doTranspose, setup):
    """Common front part of both solveLinearSystem! methods (MGWrapper.jl:28-63, SAAMGWrapper.jl:28-62).
    Returns None when B == 0 (X zeroed), else (B, nrhs)."""
    if sp.issparse(B):
        B = B.toarray()
    B = np.asarray(B)
    if B.ndim == 2 and B.shape[1] == 1:
        B = B.reshape(-1)
    if param.doClear == 1:
        mgdef.clear(param.MG)
    if np.linalg.norm(B) == 0.0:
        X[...] = 0.0
        return None
    nrhs = 1 if B.ndim == 1 else B.shape[1]
    if not hierarchyExists(param.MG):
        doTransposeIterative = (doTranspose + 1) % 2 if param.isTranspose else doTranspose
        if param.sym != 1 and doTransposeIterative == 0:
            # the hierarchy stores adjoints (SpMatMul uses Ac_mul_B); a plain A has to be transposed once
            A = _adjoint(A)
        t0 = time.perf_counter()
        setup(A, nrhs)
        param.timeSetup += time.perf_counter() - t0
        param.MG.doTranspose = doTranspose
    if param.sym != 1 and doTranspose != param.MG.doTranspose:
        t0 = time.perf_counter()
        transposeHierarchy(param.MG)
        param.timeSetup += time.perf_counter() - t0
    return B, nrhs


def solveLinearSystem(A, B, X, param, doTranspose: int = 0):
    """solveLinearSystem!(A,B,X,param,doTranspose) for MGsolver (MGWrapper.jl:27-86) and SA_AMGsolver
    (SAAMGWrapper.jl:27-80).  Returns (X, param)."""
    verbose = param.out > 0
    if isinstance(param, MGsolver):
        def setup(AT, nrhs):
            MGsetup(AT, param.MG.Meshes[0] if param.MG.Meshes else param.Mesh, param.MG, nrhs, verbose)
    elif isinstance(param, SA_AMGsolver):
        def setup(AT, nrhs):
            SA_AMGsetup(AT, param.MG, param.sym == 1, nrhs, verbose)
    else:
        raise TypeError("solveLinearSystem: param must be an MGsolver or an SA_AMGsolver")
    prep = _prepare(A, B, X, param, doTranspose, setup)
    if prep is None:
        return X, param
    B, nrhs = prep
    MG = param.MG
    t0 = time.perf_counter()
    AT1 = MG.As[0]
    if param.Krylov == "BiCGSTAB":
        X, _, num_iter, _ = solveBiCGSTAB_MG(AT1, MG, B, X, verbose)
    elif param.Krylov == "GMRES" and isinstance(param, MGsolver):
        X, _, num_iter, _ = solveGMRES_MG(AT1, MG, B, X, True, 5, verbose)
    elif param.Krylov == "PCG":
        X, _, num_iter = solveCG_MG(AT1, MG, B, X, verbose)
    elif isinstance(param, MGsolver):
        X, _, num_iter = solveMG(MG, B, X, verbose)
    else:
        raise ValueError(f"SA_AMGsolver: unknown Krylov method {param.Krylov!r}")
    param.nIter += num_iter * nrhs
    param.timeSolve += time.perf_counter() - t0
    if isinstance(param, SA_AMGsolver) and num_iter >= MG.maxOuterIter - 1:
        warnings.warn("MG solver reached maximum iterations without convergence")
    return X, param


def setupSolver(AT, s: MGsolver) -> MGsolver:
    """MGWrapper.jl:88-91."""
    MGsetup(AT, s.MG.Meshes[0] if s.MG.Meshes else s.Mesh, s.MG, 1, s.out > 0)
    return s


def copySolver(s):
    """MGWrapper.jl:94-97, SAAMGWrapper.jl:84-87: copies what is necessary, counters reset."""
    MG = mgdef.copySolver(s.MG)
    if isinstance(s, MGsolver):
        MG.Meshes = list(s.MG.Meshes)
        return MGsolver(MG, s.Krylov, s.sym, s.out, s.isTranspose, s.doClear, s.tol, 0, 0.0, 0.0, s.Mesh)
    return SA_AMGsolver(MG, s.Krylov, s.sym, s.out, s.isTranspose, s.doClear, s.tol, 0, 0.0, 0.0)


def clear(s):
    """MGWrapper.jl:101-104, SAAMGWrapper.jl:91-94."""
    mgdef.clear(s.MG)
    s.doClear = 0
