"""Solve-phase front ends with the reference's signatures (src/Multigrid/SolveFuncs.jl).

All of them run on the device through the C ABI; nothing here computes on the
CPU beyond argument checks.
"""
from __future__ import annotations

import numpy as np

from .device import uploadHierarchy
from .mgdef import MGparam, hierarchyExists
from .mgsetup import adjustMemoryForNumRHS


def _nrhs(b):
    b = np.asarray(b)
    return 1 if b.ndim == 1 else b.shape[1]


def _device(param: MGparam, b):
    if not hierarchyExists(param):
        raise RuntimeError("You have to do a setup first.")
    adjustMemoryForNumRHS(param, _nrhs(b))
    dev = uploadHierarchy(param)
    dev.set_cycle(param)
    return dev


def _is_single(dt) -> bool:
    return np.dtype(dt) in (np.dtype(np.float32), np.dtype(np.complex64))


def _krylov_device(param: MGparam, AT, b):
    """Device handle the Krylov drivers run on.  A single-precision hierarchy under double-precision b / x0 is the
    reference's mixed-precision mode (getMultigridPreconditioner, SolveFuncs.jl:52-60: the cycle runs in single
    precision on a rounded copy of the residual, the Krylov method and its matrix stay in double precision)."""
    dev = _device(param, b)
    if not (_is_single(param.VAL) and not _is_single(np.asarray(b).dtype)):
        _krylov_matrix(dev, AT, param)
        return dev
    import scipy.sparse as sp
    key = None if AT is None else id(AT)
    mixed = getattr(param, "_mixed_device", None)
    if mixed is None or mixed.inner is not dev or getattr(param, "_mixed_key", None) != key:
        if mixed is not None:
            mixed.destroy()
        Ak = param.As[0] if AT is None else AT
        kd = np.complex128 if np.iscomplexobj(Ak.data) or np.dtype(param.VAL).kind == "c" else np.float64
        mixed = type(dev).mixed_over(dev, sp.csc_matrix(Ak).astype(kd))
        param._mixed_device, param._mixed_key = mixed, key
    if mixed.nrhs != _nrhs(b):
        mixed.adjust_nrhs(_nrhs(b))
    return mixed


def solveMG(param: MGparam, b, x, verbose: bool = False):
    """solveMG(param,b,x,verbose) -> (x, param, iter)   (SolveFuncs.jl:3-39).
    ``param.last_resvec`` holds [res_init, res_1, ...] (the per-cycle residual norms)."""
    dev = _device(param, b)
    xx, it, res = dev.solveMG(b, x, param.relativeTol, param.maxOuterIter)
    param.last_resvec = res
    if verbose:
        for k in range(1, len(res)):
            print(f"Cycle {k} done with relres: {res[k] / res[0]}. Convergence factor: {res[k] / res[k - 1]}")
    x[...] = xx.reshape(x.shape)
    return x, param, it


def _same_storage(a, b):
    return (a.shape == b.shape and a.nnz == b.nnz and a.dtype == b.dtype
            and a.data.__array_interface__["data"][0] == b.data.__array_interface__["data"][0]
            and a.indices.__array_interface__["data"][0] == b.indices.__array_interface__["data"][0])


def _krylov_matrix(dev, AT, param):
    """The AT argument of solveCG_MG / solveGMRES_MG (SolveFuncs.jl:77-82): uploaded separately only
    when it is not the matrix the hierarchy was built from.  The reference builds Afun from the AT of every call
    (getAfun, SolveFuncs.jl:65-82), so a matrix left behind by an earlier solve is released here."""
    import scipy.sparse as sp
    if (AT is None or AT is param.As[0]
            or (sp.issparse(AT) and AT.format == "csc" and _same_storage(AT, param.As[0]))):
        dev.set_krylov_matrix(None)
        return
    dev.set_krylov_matrix(AT)


def solveCG_MG(AT, param: MGparam, b, x0, verbose: bool = False):
    """solveCG_MG(AT,param,b,x0,verbose) -> (x, param, iter)   (SolveFuncs.jl:77-79,103-116).
    AT is the (adjoint-stored) matrix the Krylov method multiplies with."""
    dev = _krylov_device(param, AT, b)
    xx, it, flag, res = dev.solveCG(b, x0, param.relativeTol, param.maxOuterIter)
    param.last_resvec, param.last_flag = res, flag
    x0[...] = xx.reshape(x0.shape)
    return x0, param, it


def solveBiCGSTAB_MG(AT, param: MGparam, b, x0, verbose: bool = False):
    """solveBiCGSTAB_MG(AT,param,b,x0,verbose) -> (x, param, iter, nprec)   (SolveFuncs.jl:73-75,85-99);
    an n x nrhs block b runs KrylovMethods.blockBiCGSTB (SolveFuncs.jl:95)."""
    dev = _krylov_device(param, AT, b)
    xx, it, flag, res, nprec = dev.solveBiCGSTAB(b, x0, param.relativeTol, param.maxOuterIter)
    param.last_resvec, param.last_flag = res, flag
    x0[...] = xx.reshape(x0.shape)
    return x0, param, it, nprec


def solveGMRES_MG(AT, param: MGparam, b, x0, flexible: bool, inner: int, verbose: bool = False):
    """solveGMRES_MG(AT,param,b,x0,flexible,inner,verbose) -> (x, param, iter, resvec)
    (SolveFuncs.jl:80-82,120-132)."""
    dev = _krylov_device(param, AT, b)
    xx, it, flag, res = dev.solveFGMRES(b, x0, inner, flexible, param.relativeTol, param.maxOuterIter)
    param.last_resvec, param.last_flag = res, flag
    x0[...] = xx.reshape(x0.shape)
    return x0, param, it, res


def getMultigridPreconditioner(param: MGparam, B, verbose: bool = False):
    """r -> M^-1 r: one cycle from a zero initial guess (SolveFuncs.jl:43-63)."""
    if not hierarchyExists(param):
        print("You have to do a setup first.")
    dev = _device(param, B)

    def MMG(r):
        return dev.precondition(r)
    return MMG


def recursiveCycle(param: MGparam, b, x, level: int = 1):
    """recursiveCycle(param,b,x,1) (MGcycle.jl:1-118); only level 1 is exposed."""
    if level != 1:
        raise NotImplementedError("the device path starts cycles at level 1")
    dev = _device(param, b)
    xx = dev.cycle(b, x)
    x[...] = xx.reshape(x.shape)
    return x


def SpMatMul(alpha, param: MGparam, level: int, which: str, x, beta, target):
    """SpMatMul(alpha,AT,x,beta,target) (SpMatMul.jl:4-13) on a matrix of the uploaded
    hierarchy: which in {"A","P","R"} of ``level`` (1-based)."""
    dev = uploadHierarchy(param)
    y = dev.spmatmul(level, {"A": 0, "P": 1, "R": 2}[which], alpha, x, beta, target)
    target[...] = y.reshape(target.shape)
    return target
