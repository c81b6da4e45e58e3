"""Transfer operators for systems on staggered grids (face-centred components, optional cell-centred block), mirroring
src/Multigrid/Systems.jl of the reference - SURVEY.md section 8(f) rank 4.  ``n`` is always in CELLS here
(Systems.jl:2).  Unknown ordering: [faces normal to x; faces normal to y; (faces normal to z); (cells)], every block
lexicographic with x fastest; a face block has nodes in its own direction and cells in the others.

Host setup only: MGsetup (mgsetup.py) builds ``Ps / Rs`` from these when ``param.transferOperatorType`` is
"SystemsFacesLinear" or "SystemsFacesMixedLinear" (MGsetup.jl:49-75), and the hierarchy goes through the same upload and
device cycle as every other one (general CSR path).  The Vanka smoothers that usually accompany these operators in the
reference are out of scope; the diagonal smoothers work on them unchanged.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def _csc(M):
    M = sp.csc_matrix(M)
    M.eliminate_zeros()
    M.sort_indices()
    return M


def speye(n):
    return sp.identity(int(n), format="csc")


def _check_even(n, nc, who):
    if 2 * nc != n:
        raise ValueError(f"Err: {who}(): size should be a multiplication of 2")


def get1DNodeInjection(n_cells: int):
    """Systems.jl:80-93: node injection C,F,C,F,...,C."""
    n = int(n_cells)
    if n < 8:
        return speye(n + 1), n
    nc = n // 2
    _check_even(n, nc, "get1DNodeInjection")
    return _csc(speye(n + 1)[0:n + 1:2, :]), nc


def get1DNodeFullWeightRestriction(n_cells: int):
    """Systems.jl:95-111: full weighting x 2 on nodes (0.5, 1, 0.5; the boundary rows are what the sliced tridiagonal
    leaves: 1, 0.5)."""
    n = int(n_cells)
    if n < 8:
        return speye(n + 1), n
    nc = n // 2
    _check_even(n, nc, "get1DNodeFullWeightRestriction")
    R = sp.diags([np.full(n, .25), np.full(n + 1, .5), np.full(n, .25)], [-1, 0, 1], format="csc")
    R = sp.csc_matrix(R[:, 0::2].T)
    return _csc(R * 2), nc


def get1DProlongationCellCentered(ncells_fine: int):
    """Systems.jl:114-133: two coarse cells [C, C] into [F, F, F, F] with weights 1/4, 3/4; first and last row set to 1.
    (The main diagonal of the reference's spdiagm has n - 1 entries: its last element is absent.)"""
    n = int(ncells_fine)
    if n < 8:
        return speye(n), n
    nc = n // 2
    _check_even(n, nc, "get1DProlongationCellCentered")
    P = sp.lil_matrix((n, n))
    i = np.arange(n)
    P[i[2:], i[:-2]] = .25            # diagonal -2, n - 2 entries
    P[i[1:], i[:-1]] = .75            # diagonal -1, n - 1 entries
    P[i[:-1], i[:-1]] = .75           # diagonal 0, n - 1 entries
    P[i[:-1], i[1:]] = .25            # diagonal +1, n - 1 entries
    P = P.tocsc()[:, 0::2].tolil()
    P[0, 0] = 1.0
    P[n - 1, P.shape[1] - 1] = 1.0
    return _csc(P), nc


def get1DRestrictionCells(n: int):
    """Systems.jl:135-149: 2 x 1 aggregation (weights 1, 1)."""
    n = int(n)
    if n < 8:
        return speye(n), n
    nc = n // 2
    _check_even(n, nc, "get1DRestrictionCells")
    R = sp.diags([np.full(n - 1, .5), np.full(n - 1, .5)], [0, 1], shape=(n - 1, n), format="csc")
    return _csc(2 * R[0:n:2, :]), nc


def get1DProlongationNodes(ncells_fine: int):
    """Systems.jl:151-164: linear interpolation on nodes."""
    n = int(ncells_fine)
    if n < 8:
        return speye(n + 1), n
    nc = n // 2
    _check_even(n, nc, "get1DProlongationNodes")
    half = 0.5 * np.ones(n)
    P = sp.diags([half, np.ones(n + 1), half], [-1, 0, 1], format="csc")
    return _csc(P[:, 0::2]), nc


def _kron_all(ops, who):
    if len(ops) == 3:
        return _csc(sp.kron(ops[2], sp.kron(ops[1], ops[0], format="csc"), format="csc"))
    if len(ops) == 2:
        return _csc(sp.kron(ops[1], ops[0], format="csc"))
    raise ValueError(f"{who}() : Dimension not supported!")


def _per_dim(n, j, own, other, who):
    ops, nc = [], []
    for kk in range(len(n)):
        M, c = (own if kk + 1 == j else other)(int(n[kk]))
        ops.append(M)
        nc.append(c)
    return _kron_all(ops, who), np.asarray(nc, dtype=np.int64)


def getRestrictionCellCentered(n):
    """Systems.jl:167-184."""
    return _per_dim(n, 0, get1DRestrictionCells, get1DRestrictionCells, "getRestrictionCells")


def getRestrictionFacesInjectionUj(n, j: int):
    """Systems.jl:187-206 (j is 1-based: the direction the faces are normal to)."""
    return _per_dim(n, j, get1DNodeInjection, get1DRestrictionCells, "getRestrictionFacesInjectionUj")


def getRestrictionFacesFullWeightUj(n, j: int):
    """Systems.jl:208-227."""
    return _per_dim(n, j, get1DNodeFullWeightRestriction, get1DRestrictionCells, "getRestrictionFacesFullWeightUj")


def getLinearInterpolationFacesUj(n, j: int):
    """Systems.jl:229-248."""
    return _per_dim(n, j, get1DProlongationNodes, get1DProlongationCellCentered, "getLinearInterpolationFacesUj")


def getLinearInterpolationCellCentered(n):
    """Systems.jl:250-267."""
    return _per_dim(n, 0, get1DProlongationCellCentered, get1DProlongationCellCentered,
                    "getLinearInterpolationCellCentered")


def getInjectionOperatorsSystemsFaces(n, withCellsBlock: bool):
    """Systems.jl:8-31."""
    dim = len(n)
    blocks = [getRestrictionFacesInjectionUj(n, j)[0] for j in range(1, dim + 1)]
    if withCellsBlock:
        blocks.append(getRestrictionCellCentered(n)[0])
    return _csc(sp.block_diag(blocks, format="csc"))


def getLinearOperatorsSystemsFaces(n, withCellsBlock: bool):
    """Systems.jl:33-76: (P, R, nc) - block-diagonal prolongation (linear on the nodes of a face block's own direction,
    cell-centred in the others) and full-weighting restriction, nc = coarse cells per dimension."""
    dim = len(n)
    if dim not in (2, 3):
        raise ValueError("getLinearOperatorsSystemsFaces(): Dimension not supported!")
    P1, nc = getLinearInterpolationFacesUj(n, 1)
    Pb = [P1] + [getLinearInterpolationFacesUj(n, j)[0] for j in range(2, dim + 1)]
    Rb = [getRestrictionFacesFullWeightUj(n, j)[0] for j in range(1, dim + 1)]
    if withCellsBlock:
        Pb.append(getLinearInterpolationCellCentered(n)[0])
        Rb.append(getRestrictionCellCentered(n)[0])
    return _csc(sp.block_diag(Pb, format="csc")), _csc(sp.block_diag(Rb, format="csc")), nc


def faces_size(n, withCellsBlock: bool = False) -> int:
    """Number of unknowns of a staggered system on n cells."""
    n = [int(v) for v in n]
    tot = 0
    for j in range(len(n)):
        tot += int(np.prod([n[k] + (1 if k == j else 0) for k in range(len(n))]))
    return tot + (int(np.prod(n)) if withCellsBlock else 0)
