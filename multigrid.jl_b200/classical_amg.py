"""Classical (Ruge-Stueben) AMG setup on the host, mirroring src/Multigrid/ClassicalAMG.jl (+ coloring.jl,
interpolation.jl) of the reference - SURVEY.md section 8(f) rank 4: the hierarchy it produces (``As / Ps / Rs /
relaxPrecs``) goes through the same upload and the same device cycle as the geometric and SA-AMG ones.

What is and is not reproducible.  The first colouring pass (coloring.jl:13-97) takes "a node that has a maximal
neighbour count" with ``pop!(::Set{Int})``: which of several nodes with the same count comes out is whatever slot order
Julia's hash table has at that moment.  That choice cannot be restated without Julia's ``Dict`` internals, so this
module takes the SMALLEST such node - a valid run of the same algorithm, deterministic, but not necessarily the same
C/F splitting as a Julia run.  Everything else - strength matrix, count updates, second pass, direct interpolation, the
Galerkin products, the 1e-8 shift of the coarsest matrix - follows the cited lines literally, quirks included
(e.g. coloring.jl:33-36 compares where it meant to assign, so nodes with only a diagonal entry keep count 1 and end up
coarse).  Parity is therefore "unpinned" in the sense of DESIGN.md section 3: the tests hold the reference's own
thresholds (test/Multigrid/testSAforDivSigGrad.jl:66-76,112-124) on the oracle and compare the device with the oracle
on the hierarchies this module produces.
"""
from __future__ import annotations

import heapq

import numpy as np
import scipy.sparse as sp

from .mgdef import MGparam
from .mgsetup import _csc, _invalidate_device, adjustMemoryForNumRHS, defineCoarsestAinv, galerkin, getRelaxPrec
from .sa_amg import _entry_norm, _op_norm


def getStrengthMatrixClassical(AT, strengthConnParam: float, n: int):
    """ClassicalAMG.jl:87-116: S = -AT scaled per column by its largest entry, entries below the threshold zeroed,
    diagonal set to one, then symmetrised and zeros dropped.  (Unlike SA-AMG.jl:90-114 the threshold is applied BEFORE
    the diagonal is set.)"""
    S = _csc(AT).copy()
    S.data = -S.data
    cp, rv, nz = S.indptr, S.indices, S.data
    mm = 1e-16 * nz.max()
    for j in range(n):
        a, b = cp[j], cp[j + 1]
        if b == a:
            continue
        seg = nz[a:b]
        maxVal = max(mm, seg.max())
        seg *= 1.0 / maxVal
        seg[seg < strengthConnParam] = 0.0
        seg[rv[a:b] == j] = 1.0
    St = _csc(S.T)
    return _half_sum(S, St)


def _half_sum(S, St):
    # (S + S') / 2 with dropzeros! (ClassicalAMG.jl:114-115): the sum first, then the division
    C = sp.csc_matrix(S + St)
    C.data = C.data / 2
    C.eliminate_zeros()
    C.sort_indices()
    return C


def getColoringFirst(S, n: int):
    """coloring.jl:13-97: maximal independent set of coarse nodes by strong-connection counts.  C = 1, F = 0.
    ``pop!(indeces[nmax + 1])`` of the reference is replaced by "the smallest node of that bucket" (module docstring);
    buckets are heaps with lazy deletion plus exact member counts."""
    cp, rows = S.indptr, S.indices
    lam = (cp[1:n + 1] - cp[:n]).astype(np.int64)
    coloring = np.zeros(n, dtype=np.int64)
    cap = int(lam.max()) + n + 2 if n else 2
    heaps: dict[int, list] = {}
    count = np.zeros(cap, dtype=np.int64)

    def push(v, i):
        heaps.setdefault(v, [])
        heapq.heappush(heaps[v], i)
        count[v] += 1

    nmax = 0
    for i in range(n):
        # coloring.jl:33-36: `lambda[i] == 0` / `coloring[i] == 0` are comparisons, not assignments: nothing happens
        push(int(lam[i]), i)
        if lam[i] > nmax:
            nmax = int(lam[i])
    alive = np.ones(n, dtype=bool)          # False once a node was popped as coarse (it is in no bucket afterwards)
    while nmax > 0:
        old_max = nmax
        h = heaps[nmax]
        while True:                         # lazy deletion: skip entries whose node moved to another bucket
            curr = heapq.heappop(h)
            if alive[curr] and lam[curr] == nmax:
                break
        count[nmax] -= 1
        alive[curr] = False
        coloring[curr] = 1
        lam[curr] = 0
        for jj in range(cp[curr], cp[curr + 1]):
            row = rows[jj]
            if lam[row] != 0:
                count[lam[row]] -= 1
                lam[row] = 0
                push(0, row)
                coloring[row] = 0
        for jj in range(cp[curr], cp[curr + 1]):
            row = rows[jj]
            for kk in range(cp[row], cp[row + 1]):
                rowk = rows[kk]
                if lam[rowk] == 0:
                    continue
                count[lam[rowk]] -= 1
                lam[rowk] += 1
                push(int(lam[rowk]), rowk)
                if lam[rowk] > nmax:
                    nmax = int(lam[rowk])
        if old_max == nmax:
            for j in range(nmax, -1, -1):
                nmax = j
                if count[j] > 0:
                    break
    return coloring


def getColoringSecond(S, coloring, n: int):
    """coloring.jl:104-157: every strongly connected F-F pair must share a strongly connected C node, else the current
    node becomes C.  Sequential, in place (later nodes see earlier changes)."""
    cp, rows = S.indptr, S.indices
    for i in range(n):
        if coloring[i] == 1:
            continue
        nb = rows[cp[i]:cp[i + 1]]
        nb = nb[nb != i]
        fconn = nb[coloring[nb] == 0]
        cconn = set(int(v) for v in nb[coloring[nb] == 1])
        for j in fconn:
            common = False
            for k in range(cp[j], cp[j + 1]):
                r = rows[k]
                if r == i:
                    continue
                if coloring[r] == 1 and int(r) in cconn:
                    common = True
                    break
            if not common:
                coloring[i] = 1
                break
    return coloring


def getInterpolation(AT, S, coloring, n: int):
    """interpolation.jl:3-17 with getInterpolation1 (:19-35) and getDirectInterpolation2 (:44-97): direct interpolation
    (PyAMG's formula).  Returns (P, PT) with P n x nc, as the reference does (``R`` there is PT)."""
    AT = _csc(AT)
    ones = S.copy()
    ones.data[:] = 1.0
    Sv = _csc(AT.multiply(ones))            # S .= AT .* S : AT's values on S's pattern
    Sv.sort_indices()
    cp, rv, sv = Sv.indptr, Sv.indices, Sv.data
    acp, arv, av = AT.indptr, AT.indices, AT.data
    Pp = np.zeros(n + 1, dtype=np.int64)
    nz = 0
    for i in range(n):
        if coloring[i] == 1:
            nz += 1
        else:
            seg = rv[cp[i]:cp[i + 1]]
            nz += int(np.count_nonzero((seg != i) & (coloring[seg] == 1)))
        Pp[i + 1] = nz
    Px = np.zeros(nz, dtype=np.float64)
    Pj = np.zeros(nz, dtype=np.int64)
    with np.errstate(divide="ignore", invalid="ignore"):
        for i in range(n):
            if coloring[i] == 1:
                Pj[Pp[i]] = i
                Px[Pp[i]] = 1.0
                continue
            seg, val = rv[cp[i]:cp[i + 1]], sv[cp[i]:cp[i + 1]]
            strongC = (coloring[seg] == 1) & (seg != i)
            # getStrongSum (:116-129), sequential sums in stored order
            sum_strong_pos = sum_strong_neg = 0.0
            for v in val[strongC]:
                if v > 0:
                    sum_strong_pos += v
                else:
                    sum_strong_neg += v
            # getAllSum (:132-148)
            sum_all_pos = sum_all_neg = a_ii = 0.0
            for r, v in zip(arv[acp[i]:acp[i + 1]], av[acp[i]:acp[i + 1]]):
                if r == i:
                    a_ii += v
                elif v < 0:
                    sum_all_neg += v
                else:
                    sum_all_pos += v
            i_alpha = np.float64(sum_all_neg) / np.float64(sum_strong_neg)
            i_beta = np.float64(sum_all_pos) / np.float64(sum_strong_pos)
            if sum_strong_pos == 0:
                a_ii += sum_all_pos
                i_beta = 0.0
            neg = -1 * i_alpha / np.float64(a_ii)
            pos = -1 * i_beta / np.float64(a_ii)
            k = Pp[i]
            for r, v in zip(seg[strongC], val[strongC]):
                Pj[k] = r
                Px[k] = pos * v if v > 0 else neg * v
                k += 1
    sum_map = np.concatenate(([0], np.cumsum(coloring[:n])[:-1])) if n else np.zeros(0, dtype=np.int64)
    Pj = sum_map[Pj]
    nc = int(Pj.max()) + 1 if nz else 0
    PT = sp.csc_matrix((Px, Pj, Pp), shape=(nc, n))
    PT.sort_indices()
    P = _csc(PT.T)
    return P, PT


def ClassicalAMGsetup(AT, param: MGparam, symm: bool = True, nrhs: int = 1, verbose: bool = False,
                      opnorm: bool = False):
    """ClassicalAMG.jl:5-82.  ``opnorm`` as in SA_AMGsetup (the 1e-8 shift of the coarsest matrix uses ``norm(A, 1)``,
    entrywise in the pinned Julia 1.7)."""
    if not symm:
        raise RuntimeError("not supported yet...")
    VAL = param.VAL
    if np.iscomplexobj(np.zeros(0, dtype=VAL)):
        raise TypeError("Classical AMG is real-only in the reference (Float64 arrays in interpolation.jl:45)")
    norm = _op_norm if opnorm else _entry_norm
    rVAL = np.zeros(0, dtype=VAL).real.dtype
    As, Ps, Rs, relaxPrecs = [_csc(AT, dtype=VAL)], [], [], []
    N = As[0].shape[1]
    levels = param.levels
    for l in range(levels - 1):
        ATl = As[l]
        if param.relaxType not in ("Jac", "Jac-GMRES", "SPAI"):
            raise ValueError("Unknown relaxation type !!!!")
        d = getRelaxPrec(ATl, param.relaxType, param.relaxParam, VAL)
        S = getStrengthMatrixClassical(ATl, param.strongConnParam, N)
        coloring = getColoringFirst(S, N)
        coloring = getColoringSecond(S, coloring, N)
        P, PT = getInterpolation(ATl, S, coloring, N)
        Nc = P.shape[1]
        if P.shape[0] == P.shape[1]:
            if verbose:
                print(f"Stopped Coarsening at level {l + 1}")
            levels = l + 1
            break
        relaxPrecs.append(d)
        # "we hold the transpose of the matrices and P = R' anyway here" (ClassicalAMG.jl:55-56)
        Rs.append(_csc(P, dtype=rVAL))
        Ps.append(_csc(PT, dtype=rVAL))
        As.append(_csc(galerkin(Ps[l], ATl, Rs[l]), dtype=VAL))
        if verbose:
            print(f"MG setup: {N} -> {Nc}")
        N = Nc
    if verbose:
        print("MG Setup: Operator complexity = ", sum(a.nnz for a in As) / As[0].nnz)
    nL = As[-1].shape[1]
    As[-1] = _csc(As[-1] + 1e-8 * norm(As[-1], 1) * sp.identity(nL, format="csc"), dtype=VAL)
    param.levels = levels
    param.As, param.Ps, param.Rs = As, Ps, Rs
    param.relaxPrecs = relaxPrecs
    param.Meshes = []
    defineCoarsestAinv(param, As[-1])
    _invalidate_device(param)
    adjustMemoryForNumRHS(param, nrhs, verbose)
    return None
