"""TEST INFRASTRUCTURE - literal, loop-by-loop restatement of the reference's SA-AMG
aggregation (src/Multigrid/SA-AMG.jl:88-224) in pure Python with 1-based indices, used only
by the tests to check the product's accelerated implementation bit for bit on small cases.
PARITY UNPINNED against a real Julia run (none is possible here); see oracle/cycle.py."""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def getStrengthMatrix(AT, theta):
    """SA-AMG.jl:88-116 on 1-based arrays."""
    S = sp.csc_matrix(AT).copy()
    S.sort_indices()
    colptr = S.indptr.astype(np.int64) + 1
    rowval = S.indices.astype(np.int64) + 1
    nzval = -S.data.astype(np.float64)
    mm = 1e-16 * nzval.max()
    n = S.shape[1]
    for j in range(1, n + 1):
        maxVal_j = mm
        for g in range(colptr[j - 1], colptr[j]):
            if nzval[g - 1] > maxVal_j:
                maxVal_j = nzval[g - 1]
        scal_k = 1.0 / maxVal_j
        for g in range(colptr[j - 1], colptr[j]):
            nzval[g - 1] *= scal_k
        for g in range(colptr[j - 1], colptr[j]):
            if rowval[g - 1] == j:
                nzval[g - 1] = 1.0
        for g in range(colptr[j - 1], colptr[j]):
            if nzval[g - 1] < theta:
                nzval[g - 1] = 0.0
    S = sp.csc_matrix((nzval, rowval - 1, colptr - 1), shape=S.shape)
    S = sp.csc_matrix(S + S.T)      # Julia's sparse + does not store numerically-zero results
    S.eliminate_zeros()
    S.sort_indices()
    return S


def neighborhoodAggregationNew(S):
    """SA-AMG.jl:119-211, variable for variable."""
    S = sp.csc_matrix(S)
    S.sort_indices()
    colptr = (S.indptr.astype(np.int64) + 1).tolist()
    rowval = (S.indices.astype(np.int64) + 1).tolist()
    nzval = S.data.astype(np.float64).tolist()
    tau = 3.0
    n = S.shape[1]
    aggr = [0] * (n + 1)          # 1-based, aggr[0] unused
    aux = [0.0] * (n + 1)
    aux_count = [0] * (n + 1)
    avg_sparsity = 0.0
    for k in range(1, n + 1):
        avg_sparsity += colptr[k] - colptr[k - 1]
    avg_sparsity /= n
    for k in range(1, n + 1):
        if colptr[k] - colptr[k - 1] > tau * avg_sparsity:
            aux_count[k] = -1
    for k in range(1, n + 1):
        flag = False
        if aux_count[k] == -1:
            continue
        for g in range(colptr[k - 1], colptr[k]):
            if aggr[rowval[g - 1]] != 0:
                flag = True
                break
        if not flag:
            for g in range(colptr[k - 1], colptr[k]):
                if aux_count[rowval[g - 1]] != -1:
                    aggr[rowval[g - 1]] = k
                    aux_count[k] += 1
    for k in range(1, n + 1):
        flag = False
        if aux_count[k] != -1:
            continue
        aux_count[k] = 0
        for g in range(colptr[k - 1], colptr[k]):
            if aggr[rowval[g - 1]] != 0:
                flag = True
                break
        if not flag:
            for g in range(colptr[k - 1], colptr[k]):
                aggr[rowval[g - 1]] = k
                aux_count[k] += 1
    for k in range(1, n + 1):
        chosen_score = 0.0
        chosen = 0
        if aggr[k] == 0:
            for g in range(colptr[k - 1], colptr[k]):
                if aggr[rowval[g - 1]] > 0:
                    agg_of_neighbor = aggr[rowval[g - 1]]
                    aux[agg_of_neighbor] += nzval[g - 1]
                for g2 in range(colptr[k - 1], colptr[k]):
                    if aggr[rowval[g2 - 1]] > 0:
                        agg_of_neighbor = aggr[rowval[g2 - 1]]
                        if chosen_score < aux[agg_of_neighbor] / aux_count[agg_of_neighbor]:
                            chosen_score = aux[agg_of_neighbor] / aux_count[agg_of_neighbor]
                            chosen = agg_of_neighbor
                            aux[agg_of_neighbor] = 0
                aggr[k] = -chosen
    for k in range(1, n + 1):
        if aggr[k] < 0:
            aggr[k] = -aggr[k]
    return np.asarray(aggr[1:], dtype=np.int64)
