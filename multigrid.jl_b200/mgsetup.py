"""Host-side geometric setup, mirroring src/Multigrid/MGsetup.jl.

The setup stays on the host (BASELINE.json north_star: "the hierarchy is still
built by the reference's host-side setup ... and uploaded once as device CSR
arrays").  Everything produced here is in the reference's storage convention
(adjoint CSC, see mgdef.py); ``device.upload_hierarchy`` hands the raw arrays
to the C ABI.
"""
from __future__ import annotations

import copy as _copy

import numpy as np
import scipy.sparse as sp

from .mesh import RegularMesh, getRegularMesh
from .mgdef import MGparam, multilevelOperatorConstructor
from .transfer import getFWInterp


def _csc(M, dtype=None):
    M = sp.csc_matrix(M) if dtype is None else sp.csc_matrix(M, dtype=dtype)
    if not M.has_sorted_indices:
        M.sort_indices()
    return M


def _adjoint_csc(M):
    """sparse(M') of the reference: conjugate transpose, CSC, sorted rows."""
    return _csc(sp.csc_matrix(M).conj().T)


def galerkin(PT, AT, RT):
    """``Ps[l]*AT*Rs[l]`` (MGsetup.jl:102), evaluated left to right as Julia does."""
    return _csc((PT @ AT) @ RT)


def getSPAIprec(AT):
    """MGsetup.jl:359-362: conj(diag(AT)) ./ rowsumsq(AT)."""
    AT = sp.csr_matrix(AT)
    s = np.asarray(AT.multiply(AT.conj()).real.sum(axis=1)).ravel()
    return np.conj(AT.diagonal()) / s


def getRelaxPrec(AT, relaxType, relaxParam=1.0, VAL=None):
    """MGsetup.jl:142-160 for the diagonal smoothers ("Jac", "Jac-GMRES", "SPAI")."""
    VAL = AT.dtype if VAL is None else VAL
    if relaxType in ("Jac", "Jac-GMRES"):
        d = np.conj(relaxParam / AT.diagonal())
    elif relaxType == "SPAI":
        d = np.conj(relaxParam * getSPAIprec(AT))
    else:
        raise ValueError("Unknown relaxation type !!!!")
    return np.ascontiguousarray(d, dtype=VAL)


def _relax_param_array(param: MGparam):
    if isinstance(param.relaxParam, (list, tuple, np.ndarray)):
        return list(param.relaxParam)
    return [_copy.copy(param.relaxParam) for _ in range(param.levels)]


def defineCoarsestAinv(param: MGparam, AT):
    """MGsetup.jl:323-355.  The default branch (``lu(sparse(AT'))``) is replaced
    by a dense LU on the device, factorised when the hierarchy is uploaded; the
    host only records which kind of coarsest solver is requested."""
    if param.coarseSolveType == "MUMPS":
        raise NotImplementedError("MUMPS coarsest solver is dead code in the reference "
                                  "(MGcycle.jl:150-151) and is not provided")
    if param.coarseSolveType == "GMRES":
        param.LU = np.ascontiguousarray(np.conj(param.relaxParam / AT.diagonal()), dtype=param.VAL)
    elif param.coarseSolveType == "VankaFaces":
        raise NotImplementedError("VankaFaces coarsest solver is out of scope")
    else:
        param.LU = "device-dense-LU"


def MGsetup(ATf, Mesh: RegularMesh, param: MGparam, nrhs: int = 1, verbose: bool = False):
    """Geometric multigrid setup (MGsetup.jl:7-138).

    ``ATf`` is either the sparse matrix A^H (callers pass A' themselves, tests
    pass symmetric real A) or a multilevelOperatorConstructor (rediscretisation
    on every level; the adjoint is taken here, MGsetup.jl:28,106)."""
    VAL = param.VAL
    rVAL = np.zeros(0, dtype=VAL).real.dtype
    relaxParamArr = _relax_param_array(param)
    geometric = isinstance(ATf, multilevelOperatorConstructor)
    PDEparam = 0
    if not geometric:
        A1 = _csc(ATf)
    else:
        A1 = _adjoint_csc(ATf.getOperator(Mesh, ATf.param))
        PDEparam = ATf.param
    if A1.dtype != VAL:
        A1 = _csc(A1, dtype=VAL)
    if param.transferOperatorType not in ("FullWeighting", "SystemsFacesLinear", "SystemsFacesMixedLinear"):
        raise ValueError(f"unknown transferOperatorType {param.transferOperatorType!r}")
    withCellsBlock = param.transferOperatorType == "SystemsFacesMixedLinear"      # MGsetup.jl:48-51
    As, Ps, Rs, Meshes, relaxPrecs = [A1], [], [], [Mesh], []
    n = Mesh.n.copy()
    Cop = A1.nnz
    levels = param.levels
    for l in range(levels - 1):
        AT = As[l]
        if param.transferOperatorType == "FullWeighting":
            P, nc = getFWInterp(n + 1, geometric)
            nc = nc - 1
            RT = _csc(P.copy(), dtype=rVAL)
            PT = _csc(P.T, dtype=rVAL)
        else:
            # staggered-grid systems (MGsetup.jl:63-74, Systems.jl:33-76): n in cells, block-diagonal P and R
            from .systems import getLinearOperatorsSystemsFaces
            P, R, nc = getLinearOperatorsSystemsFaces(n, withCellsBlock)
            PT = _csc(P.T, dtype=rVAL)
            RT = _csc(R.T, dtype=rVAL)
        RT.data *= 0.5 ** Meshes[l].dim
        relaxPrecs.append(getRelaxPrec(AT, param.relaxType, relaxParamArr[l], VAL))
        if PT.shape[0] == PT.shape[1]:
            if verbose:
                print(f"Stopped Coarsening at level {l + 1}")
            # the reference keeps relaxPrecs[1:l] here (MGsetup.jl:84-92); the last
            # entry belongs to what is now the coarsest level and is never used
            levels = l + 1
            break
        Ps.append(PT)
        Rs.append(RT)
        Meshes.append(getRegularMesh(Meshes[l].domain, nc))
        if not geometric:
            Act = galerkin(PT, AT, RT)
        else:
            PDEparam = ATf.restrictParams(Meshes[l], Meshes[l + 1], PDEparam, l + 1)
            Act = _adjoint_csc(ATf.getOperator(Meshes[l + 1], PDEparam))
        if Act.dtype != VAL:
            Act = _csc(Act, dtype=VAL)
        As.append(Act)
        Cop += Act.nnz
        if verbose:
            print(f"MG setup: {n.tolist()} cells done")
        n = nc
    if verbose:
        print("MG setup: Operator complexity = ", Cop / As[0].nnz)
    param.levels = levels
    param.As = As
    param.Meshes = Meshes
    defineCoarsestAinv(param, As[-1])
    param.Ps = Ps
    param.Rs = Rs
    param.relaxPrecs = relaxPrecs[:max(levels - 1, 0)] if len(relaxPrecs) >= levels else relaxPrecs
    _invalidate_device(param)
    adjustMemoryForNumRHS(param, nrhs, verbose)
    param.doTranspose = 0
    return param


def _invalidate_device(param: MGparam):
    if getattr(param, "_mixed_device", None) is not None:
        param._mixed_device.destroy()   # before the hierarchy it preconditions with
        param._mixed_device = None
    if param.device is not None:
        param.device.destroy()
        param.device = None


def adjustMemoryForNumRHS(param: MGparam, nrhs: int = 1, verbose: bool = False):
    """MGsetup.jl:166-223.  The reference (re)allocates CYCLEmem/FGMRESmem on
    the host; here the workspaces are device buffers, resized lazily when the
    hierarchy is uploaded or nrhs changes."""
    if len(param.As) == 0:
        raise RuntimeError("The Hierarchy is empty - run a setup first.")
    param.nrhs = int(nrhs)
    if param.device is not None:
        param.device.adjust_nrhs(int(nrhs))
    return param


def replaceMatrixInHierarchy(param: MGparam, AT, verbose: bool = False, device: bool = True):
    """MGsetup.jl:226-270: keep Ps/Rs, redo the Galerkin products, the
    relaxation diagonals and the coarsest factorisation.

    With a resident device hierarchy (``param.device``) and ``device=True`` the products run on the device
    (``mgb200_replace_matrix``, csrc/galerkin.cuh): only ``As[1]`` travels; ``param.As[2:]`` and ``param.relaxPrecs`` on
    the host are then refreshed from the device (values only - the sparsity is the one of the first setup).  When
    the device path does not apply (another sparsity, row-partitioned hierarchy) the host path below runs and the
    device hierarchy is dropped for re-upload, as before."""
    relaxParamArr = _relax_param_array(param)
    dev = getattr(param, "device", None)
    if (device and dev is not None and hasattr(dev, "replace_matrix") and getattr(param, "_mixed_device", None) is None
            and param.coarseSolveType not in ("MUMPS", "VankaFaces")):
        ATc = _csc(AT, dtype=param.VAL)
        rp = np.array([float(np.real(v)) for v in relaxParamArr], dtype=np.float64)
        if dev.replace_matrix(ATc, param.relaxType, rp):
            param.As[0] = ATc
            for l in range(param.levels - 1):
                param.relaxPrecs[l] = dev.download_relax_prec(l + 1, param.As[l].shape[1])
                nxt = sp.csc_matrix(param.As[l + 1], copy=True)
                if not nxt.has_sorted_indices:
                    nxt.sort_indices()
                nxt.data[:] = dev.download_values(l + 2, 0, nxt.nnz)
                param.As[l + 1] = nxt
            defineCoarsestAinv(param, param.As[-1])
            param.doTranspose = 0
            return
    param.As[0] = _csc(AT, dtype=param.VAL)
    for l in range(param.levels - 1):
        ATl = param.As[l]
        param.relaxPrecs[l] = getRelaxPrec(ATl, param.relaxType, relaxParamArr[l], param.VAL)
        param.As[l + 1] = _csc(galerkin(param.Ps[l], ATl, param.Rs[l]), dtype=param.VAL)
    defineCoarsestAinv(param, param.As[-1])
    param.doTranspose = 0
    _invalidate_device(param)


def transposeHierarchy(param: MGparam, verbose: bool = False):
    """MGsetup.jl:274-318: the hierarchy of A becomes the hierarchy of A^H.
    (The reference assigns ``Ps[l] = sparse(Rs[l]')`` and then
    ``Rs[l] = sparse(Ps[l]')`` from the already overwritten Ps, i.e. Rs is left
    unchanged and Ps becomes Rs'; this is restated literally.)"""
    if param.relaxType not in ("Jac", "Jac-GMRES", "SPAI"):
        raise RuntimeError("Not supported")
    param.As[0] = _adjoint_csc(param.As[0])
    param.doTranspose = (param.doTranspose + 1) % 2
    for l in range(param.levels - 1):
        param.relaxPrecs[l] = np.conj(param.relaxPrecs[l])
        param.Ps[l] = _adjoint_csc(param.Rs[l])
        param.Rs[l] = _adjoint_csc(param.Ps[l])
        param.As[l + 1] = _adjoint_csc(param.As[l + 1])
    if param.coarseSolveType in ("BiCGSTAB", "GMRES"):
        param.LU = np.conj(param.LU)
    _invalidate_device(param)
