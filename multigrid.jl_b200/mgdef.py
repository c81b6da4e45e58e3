"""MGparam and its constructor, mirroring src/Multigrid/MGdef.jl:91-211.

The hierarchy fields keep the reference's storage convention
(MGdef.jl:75-77): ``As[l]``, ``Ps[l]``, ``Rs[l]`` are CSC matrices of the
ADJOINT operators (A^H, P^T, R^T), so their CSC arrays are the CSR arrays of
conj(operator).  The per-level workspaces (CYCLEmem / FGMRESmem,
MGdef.jl:56-60, FGMRES.jl:3-8) live on the device in this framework; the
``device`` field holds the uploaded hierarchy handle.
"""
from __future__ import annotations

import copy as _copy

import numpy as np

Float64 = np.float64
ComplexF64 = np.complex128
Float32 = np.float32
ComplexF32 = np.complex64
Int64 = np.int64


class multilevelOperatorConstructor:
    """MGdef.jl:31-46: getOperator(mesh,param) builds the PDE on a mesh,
    restrictParams(mesh_f, mesh_c, param_f, level) gives the coarse param."""

    def __init__(self, param, getOperator, restrictParams):
        self.param = param
        self.getOperator = getOperator
        self.restrictParams = restrictParams


def getMultilevelOperatorConstructor(param, getOperator, restrictParams):
    if restrictParams == [] or restrictParams is None:
        return multilevelOperatorConstructor(
            [], lambda mesh, p: getOperator(mesh), lambda mf, mc, pf, level: [])
    return multilevelOperatorConstructor(param, getOperator, restrictParams)


class MGparam:
    """Field names follow MGdef.jl:91-116."""

    def __init__(self, VAL, IND, levels, numCores, maxOuterIter, relativeTol, relaxType,
                 relaxParam, relaxPre, relaxPost, cycleType, coarseSolveType,
                 strongConnParam, FilteringParam, transferOperatorType):
        self.VAL = np.dtype(VAL).type
        self.IND = np.dtype(IND).type
        self.levels = int(levels)
        self.numCores = int(numCores)
        self.maxOuterIter = int(maxOuterIter)
        self.relativeTol = float(relativeTol)
        self.relaxType = relaxType
        self.relaxParam = relaxParam
        self.relaxPre = relaxPre
        self.relaxPost = relaxPost
        self.cycleType = cycleType
        self.Ps = []
        self.Rs = []
        self.As = []
        self.relaxPrecs = []
        self.coarseSolveType = coarseSolveType
        self.LU = []
        self.doTranspose = 0
        self.strongConnParam = float(strongConnParam)
        self.FilteringParam = float(FilteringParam)  # stored, never read (MGdef.jl:112)
        self.Meshes = []
        self.transferOperatorType = transferOperatorType
        self.singlePrecision = self.VAL in (np.float32, np.complex64)
        self.nrhs = 0          # nrhs the device workspaces are sized for
        self.device = None     # uploaded hierarchy (multigrid_jl_b200.device.DeviceHierarchy)
        self._mixed_device = None  # double-precision Krylov handle over a single-precision hierarchy
        self._mixed_key = None
        self.aggregates = []   # SA-AMG integer maps per level (kept for parity checks)


def getMGparam(VAL=np.float64, IND=np.int64, levels=3, numCores=8, maxIter=20, relativeTol=1e-6,
               relaxType="SPAI", relaxParam=1.0, relaxPre=2, relaxPost=2, cycleType='V',
               coarseSolveType="NoMUMPS", strongConnParam=0.4, FilteringParam=0.0,
               transferOperatorType="FullWeighting"):
    """MGdef.jl:149-161; relaxPre/relaxPost may be ints or level->int functions
    (levels are 1-based as in the reference)."""
    pre = relaxPre if callable(relaxPre) else (lambda level, _v=int(relaxPre): _v)
    post = relaxPost if callable(relaxPost) else (lambda level, _v=int(relaxPost): _v)
    if cycleType not in ('V', 'F', 'W', 'K'):
        raise ValueError("cycleType must be one of 'V','F','W','K'")
    return MGparam(VAL, IND, levels, numCores, maxIter, relativeTol, relaxType, relaxParam,
                   pre, post, cycleType, coarseSolveType, strongConnParam, FilteringParam,
                   transferOperatorType)


def hierarchyExists(param: MGparam) -> bool:
    """MGdef.jl:208-210."""
    return len(param.As) > 0


def destroyCoarsestLU(param: MGparam):
    """MGdef.jl:191-206 (the factor lives on the device)."""
    param.LU = []
    if param.device is not None:
        param.device.destroy_coarsest()


def clear(param: MGparam):
    """``clear!`` (MGdef.jl:179-189): drop the hierarchy and the device copy."""
    param.Ps, param.Rs, param.As = [], [], []
    param.relaxPrecs = []
    param.Meshes = []
    param.aggregates = []
    param.LU = []
    param.nrhs = 0
    if getattr(param, "_mixed_device", None) is not None:
        param._mixed_device.destroy()   # before the hierarchy it preconditions with
        param._mixed_device = None
    if param.device is not None:
        param.device.destroy()
        param.device = None


def copySolver(MG: MGparam) -> MGparam:
    """MGdef.jl:138-145: parameters only, no hierarchy."""
    return getMGparam(MG.VAL, MG.IND, MG.levels, MG.numCores, MG.maxOuterIter, MG.relativeTol,
                      MG.relaxType, _copy.deepcopy(MG.relaxParam), MG.relaxPre, MG.relaxPost,
                      MG.cycleType, MG.coarseSolveType, MG.strongConnParam, MG.FilteringParam,
                      MG.transferOperatorType)
