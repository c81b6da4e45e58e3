"""Geometric transfer operators on nodal grids.

Behaviour follows src/Multigrid/GeometricTransferOperators.jl:5-46 of the
reference: 1-D linear interpolation on nodes, N-D by Kronecker products with x
fastest (``P = kron(P3, kron(P2, P1))``).
"""
from __future__ import annotations

import warnings

import numpy as np
import scipy.sparse as sp


def get1DFWInterp(n_nodes: int, geometric: bool):
    """1-D prolongation on ``n_nodes`` NODES (GeometricTransferOperators.jl:22-46).

    odd  : tridiag(1/2,1,1/2)[:, ::2]   (coarse nodes = odd fine nodes)
    even : geometric -> identity (that dimension stops coarsening, with a warning);
           otherwise keep the last node as an extra coarse node.
    <=2  : identity.
    """
    n_nodes = int(n_nodes)
    if n_nodes > 2:
        half = 0.5 * np.ones(n_nodes - 1)
        P = sp.diags([half, np.ones(n_nodes), half], [-1, 0, 1], format="csc")
        if n_nodes % 2 == 1:
            P = P[:, 0::2]
        elif geometric:
            P = sp.identity(n_nodes, format="csc")
            warnings.warn("getFWInterp(): in geometric mode we stop coarsening because "
                          "num cells does not divide by two")
        else:
            cols = list(range(0, n_nodes, 2)) + [n_nodes - 1]
            P = P[:, cols].tolil()
            P[n_nodes - 2:, P.shape[1] - 2:] = np.eye(2)
            P = P.tocsc()
            P.eliminate_zeros()
    else:
        P = sp.identity(n_nodes, format="csc")
    P = sp.csc_matrix(P)
    P.sort_indices()
    return P, P.shape[1]


def getFWInterp(n_nodes, geometric: bool = False):
    """Bi/tri-linear prolongation; ``n_nodes`` is the number of NODES per dim
    (GeometricTransferOperators.jl:5-20).  Returns (P, nc_nodes)."""
    n_nodes = [int(k) for k in n_nodes]
    Ps, ncs = zip(*(get1DFWInterp(k, geometric) for k in n_nodes))
    if len(n_nodes) == 2:
        P = sp.kron(Ps[1], Ps[0], format="csc")
    else:
        P = sp.kron(Ps[2], sp.kron(Ps[1], Ps[0], format="csc"), format="csc")
    P.sort_indices()
    return P, np.asarray(ncs, dtype=np.int64)
