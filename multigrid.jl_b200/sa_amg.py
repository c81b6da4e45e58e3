"""Smoothed-aggregation AMG setup on the host, mirroring src/Multigrid/SA-AMG.jl.

The integer outputs (aggregates, fine-to-coarse maps, sparsity patterns) must
be bit-exact with the reference (BASELINE.json north_star), so the greedy
aggregation is a literal restatement of SA-AMG.jl:119-211 including its
quirks; numba only compiles the same loops.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from .mgdef import MGparam
from .mgsetup import (_csc, _invalidate_device, adjustMemoryForNumRHS, defineCoarsestAinv,
                      galerkin, getRelaxPrec)

try:  # numba is only an accelerator for the literal loops below
    import numba as _nb
    _njit = _nb.njit(cache=True)
except Exception:  # pragma: no cover
    _nb = None

    def _njit(f):
        return f


def sparse_add_dropzeros(A, B):
    """Julia's sparse ``A + B`` (``map`` with zero-preserving f) does not store
    numerically-zero results; scipy's binop behaves the same, made explicit."""
    C = sp.csc_matrix(A + B)
    C.eliminate_zeros()
    C.sort_indices()
    return C


@_njit
def _strength_scale(colptr, rowval, nzval, theta):
    # SA-AMG.jl:90-114, 0-based indices
    mm = 1e-16 * nzval.max()
    n = colptr.shape[0] - 1
    for j in range(n):
        maxVal = mm
        for g in range(colptr[j], colptr[j + 1]):
            if nzval[g] > maxVal:
                maxVal = nzval[g]
        scal = 1.0 / maxVal
        for g in range(colptr[j], colptr[j + 1]):
            nzval[g] *= scal
        for g in range(colptr[j], colptr[j + 1]):
            if rowval[g] == j:
                nzval[g] = 1.0
        for g in range(colptr[j], colptr[j + 1]):
            if nzval[g] < theta:
                nzval[g] = 0.0


def getStrengthMatrix(AT, strengthConnParam: float):
    """SA-AMG.jl:88-116."""
    S = _csc(AT).copy()
    S.data = -S.data
    nz = np.ascontiguousarray(S.data, dtype=np.float64)
    _strength_scale(S.indptr.astype(np.int64), S.indices.astype(np.int64), nz, float(strengthConnParam))
    S.data = nz
    return sparse_add_dropzeros(S, sp.csc_matrix(S.T))


@_njit
def _neighborhood_aggregation(colptr, rowval, nzval):
    # literal restatement of SA-AMG.jl:119-211 with 0-based node ids; aggregate ids
    # are stored 1-based (root index + 1) so that 0 keeps meaning "unaggregated"
    # and the sign trick of pass 3 works unchanged.
    tau = 3.0
    n = colptr.shape[0] - 1
    aggr = np.zeros(n, dtype=np.int64)
    aux = np.zeros(n, dtype=np.float64)
    aux_count = np.zeros(n, dtype=np.int64)
    avg_sparsity = 0.0
    for k in range(n):
        avg_sparsity += colptr[k + 1] - colptr[k]
    avg_sparsity /= n
    for k in range(n):
        if colptr[k + 1] - colptr[k] > tau * avg_sparsity:
            aux_count[k] = -1
    # pass 1 (:139-158)
    for k in range(n):
        flag = False
        if aux_count[k] == -1:
            continue
        for g in range(colptr[k], colptr[k + 1]):
            if aggr[rowval[g]] != 0:
                flag = True
                break
        if not flag:
            for g in range(colptr[k], colptr[k + 1]):
                if aux_count[rowval[g]] != -1:
                    aggr[rowval[g]] = k + 1
                    aux_count[k] += 1
    # pass 2 (:160-178)
    for k in range(n):
        flag = False
        if aux_count[k] != -1:
            continue
        aux_count[k] = 0
        for g in range(colptr[k], colptr[k + 1]):
            if aggr[rowval[g]] != 0:
                flag = True
                break
        if not flag:
            for g in range(colptr[k], colptr[k + 1]):
                aggr[rowval[g]] = k + 1
                aux_count[k] += 1
    # pass 3 (:180-202)
    for k in range(n):
        chosen_score = 0.0
        chosen = 0
        if aggr[k] == 0:
            for g in range(colptr[k], colptr[k + 1]):
                if aggr[rowval[g]] > 0:
                    a = aggr[rowval[g]]
                    aux[a - 1] += nzval[g]
                for g2 in range(colptr[k], colptr[k + 1]):
                    if aggr[rowval[g2]] > 0:
                        a = aggr[rowval[g2]]
                        if chosen_score < aux[a - 1] / aux_count[a - 1]:
                            chosen_score = aux[a - 1] / aux_count[a - 1]
                            chosen = a
                            aux[a - 1] = 0.0
                aggr[k] = -chosen
    # pass 4 (:204-208)
    for k in range(n):
        if aggr[k] < 0:
            aggr[k] = -aggr[k]
    return aggr


def neighborhoodAggregationNew(S):
    """Returns the 1-based aggregate-root array ``aggr`` exactly as the
    reference does (SA-AMG.jl:119-211)."""
    S = _csc(S)
    return _neighborhood_aggregation(S.indptr.astype(np.int64), S.indices.astype(np.int64),
                                     np.ascontiguousarray(S.data, dtype=np.float64))


def aggrArray2P(aggr):
    """SA-AMG.jl:213-224.  Returns (P0 (n x nc, CSC), fine2coarse aggregate ids 1-based)."""
    aggr = np.asarray(aggr, dtype=np.int64)
    n = aggr.shape[0]
    if np.any(aggr <= 0):
        raise RuntimeError("nodes without aggregates")
    roots = np.nonzero(aggr == np.arange(1, n + 1))[0]
    fine2coarse = np.zeros(n, dtype=np.int64)
    fine2coarse[roots] = np.arange(1, roots.shape[0] + 1)
    agg = fine2coarse[aggr - 1]
    if np.any(agg == 0):
        raise RuntimeError("nodes without aggregates")
    P = sp.csc_matrix((np.ones(n), (np.arange(n), agg - 1)), shape=(n, roots.shape[0]))
    P.sort_indices()
    return P, agg


def getAggregation(AT, strengthConnParam: float):
    """SA-AMG.jl:78-86; n <= 100 -> identity (coarsening stops)."""
    n = AT.shape[1]
    if n <= 100:
        return sp.identity(n, format="csc"), np.arange(1, n + 1, dtype=np.int64)
    S = getStrengthMatrix(AT, strengthConnParam)
    aggr = neighborhoodAggregationNew(S)
    return aggrArray2P(aggr)


def _entry_norm(M, p):
    """Julia >= 1.0 ``norm(sparse, p)``: entrywise vector norm."""
    a = np.abs(M.data)
    if a.size == 0:
        return 0.0
    return float(a.sum()) if p == 1 else float(a.max())


def _op_norm(M, p):
    a = abs(sp.csc_matrix(M))
    return float(a.sum(axis=0).max()) if p == 1 else float(a.sum(axis=1).max())


def SA_AMGsetup(AT, param: MGparam, symm: bool = True, nrhs: int = 1, verbose: bool = False,
                opnorm: bool = False):
    """SA-AMG.jl:8-76.  ``opnorm=False`` reproduces the pinned Julia 1.7
    behaviour (``norm`` of a sparse matrix is entrywise); ``opnorm=True`` is
    the operator-norm reading the code was written for (Julia 0.6)."""
    if not symm:
        raise RuntimeError("not supported yet...")
    VAL = param.VAL
    if np.iscomplexobj(np.zeros(0, dtype=VAL)):
        raise TypeError("SA-AMG is real-only in the reference (SURVEY appendix F)")
    norm = _op_norm if opnorm else _entry_norm
    rVAL = np.zeros(0, dtype=VAL).real.dtype
    As, Ps, Rs, relaxPrecs, aggregates = [_csc(AT, dtype=VAL)], [], [], [], []
    levels = param.levels
    for l in range(levels - 1):
        ATl = As[l]
        if param.relaxType not in ("Jac", "Jac-GMRES", "SPAI"):
            raise ValueError("Unknown relaxation type !!!!")
        d = getRelaxPrec(ATl, param.relaxType, param.relaxParam, VAL)
        P0, agg = getAggregation(ATl, param.strongConnParam)
        P0T = _csc(P0.T)
        if P0T.shape[0] == P0T.shape[1]:
            if verbose:
                print(f"Stopped Coarsening at level {l + 1}")
            levels = l + 1
            break
        relaxPrecs.append(d)
        aggregates.append(agg)
        DAT = _csc(ATl @ sp.diags(d))
        rhoDAT = min(norm(DAT, 1), norm(DAT, np.inf))
        # Julia: 1.33/rhoDAT is Float64 also for a Float32 hierarchy, PT is formed in the wider type and rounded to
        # real(VAL) when it is stored into Ps / Rs (typed arrays, SA-AMG.jl:9-10,41-42); the Galerkin product then
        # runs in VAL
        PT = sparse_add_dropzeros(P0T, -(((1.33 / float(rhoDAT)) * P0T) @ DAT))
        if PT.dtype != rVAL:
            PT = _csc(PT, dtype=rVAL)
        Rs.append(_csc(PT.T))
        Ps.append(PT)
        As.append(_csc(galerkin(Ps[l], ATl, Rs[l]), dtype=VAL))
        if verbose:
            print(f"SA-AMG setup: level {l + 1}: {ATl.shape[0]} -> {PT.shape[0]}")
    if verbose:
        print("MG Setup: Operator complexity = ", sum(a.nnz for a in As) / As[0].nnz)
    nL = As[-1].shape[1]
    As[-1] = _csc(As[-1] + 1e-8 * norm(As[-1], 1) * sp.identity(nL, format="csc"), dtype=VAL)
    param.levels = levels
    param.As, param.Ps, param.Rs = As, Ps, Rs
    param.relaxPrecs = relaxPrecs
    param.aggregates = aggregates
    param.Meshes = []
    defineCoarsestAinv(param, As[-1])
    _invalidate_device(param)
    adjustMemoryForNumRHS(param, nrhs, verbose)
    return None
