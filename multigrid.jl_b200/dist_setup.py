"""Host-side setup of a z-slab row partition of a geometric hierarchy (multi-GPU path).

Layout: the reference's DomainDecomposition box partition with ``NumCells = [1,1,G]``
(src/DomainDecomposition/DDIndices.jl:41-47): ``cellSize = div(nc, NumCells)``, subdomain i
covers cells (i-1)*cellSize+1 ... i*cellSize and the last one absorbs the remainder.  With the
x-fastest lexicographic ordering every slab is a contiguous row range.  Nodal index sets of the
reference share the interface plane (DDIndices.jl:147,155); a row partition cannot, so slab g
owns node planes g*c ... (g+1)*c-1 (0-based) and the last slab also owns the final plane
(SURVEY.md section 8(e)).

Each rank builds ONLY its rows of every distributed level.  Galerkin products are computed on a
window of planes around the slab ("windowed RAP"): the fine operator restricted to the window
is a principal sub-matrix of the global operator, and a coarse row is exact as soon as its
dependency cone lies inside the window; rows near the window ends are discarded.  Levels whose
global size drops below ``replicate_below`` rows are assembled globally on every rank
(agglomeration): their part of the cycle runs redundantly with no communication.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp

from .mesh import RegularMesh, getRegularMesh, nodal_stencil_matrix
from .mgsetup import _csc, galerkin, getRelaxPrec
from .transfer import get1DFWInterp


def slab_planes(n3_cells: int, world: int):
    """Owned node-plane ranges [lo, hi) per rank for n3_cells cells in z (n3_cells+1 planes),
    following getOriginalBoundingBoxCells (DDIndices.jl:41-47) for cells and giving the last
    slab the final plane."""
    c = n3_cells // world
    if c < 1:
        raise ValueError("more ranks than cell planes")
    lo = [g * c for g in range(world)]
    hi = [(g + 1) * c for g in range(world)]
    hi[-1] = n3_cells + 1
    return list(zip(lo, hi))


def owner_ranges_for_level(n3_cells_level: int, world: int):
    return slab_planes(n3_cells_level, world)


@dataclass
class DistLevel:
    """Owned rows of one distributed level, CSR-of-adjoint arrays as scipy CSC blocks:
    ``AT`` has shape (n_global, n_owned): column j is owned row j of A (global row row_lo + j),
    row indices are GLOBAL column indices of the operator - i.e. the reference's storage
    convention restricted to the owned columns of the stored adjoint."""
    n_global: int
    row_offsets: np.ndarray          # world+1 global row offsets of this level
    AT: sp.csc_matrix                # (n_global, n_owned)
    PT: sp.csc_matrix                # (nc_global, n_owned)      rows of P owned = fine owned rows
    RT: sp.csc_matrix                # (n_global, nc_owned)      rows of R owned = coarse owned rows
    d: np.ndarray                    # relaxPrecs for the owned rows
    nc_global: int
    coarse_row_offsets: np.ndarray


@dataclass
class DistHierarchy:
    rank: int
    world: int
    dist_levels: list = field(default_factory=list)   # DistLevel for levels 1..k-1
    replicated: object = None                         # MGparam-like with As/Ps/Rs/relaxPrecs of levels k..L (global)
    n_cells: np.ndarray = None
    levels: int = 0


def poisson_window_operator(domain, n_cells, rel_shift=1e-4, kappa2=None, gamma=0.5):
    """Returns f(level_cells, plane_lo, plane_hi) -> principal sub-matrix (CSC, window-local
    indices) of the global constant-coefficient nodal operator on node planes [plane_lo,plane_hi)
    of the mesh with ``level_cells`` cells: Poisson + shift, or the shifted-Laplacian Helmholtz
    operator when kappa2 is given.  Only the fine level uses it (coarser levels are Galerkin)."""
    domain = np.asarray(domain, dtype=np.float64)

    def build(level_cells, plo, phi):
        level_cells = np.asarray(level_cells, dtype=np.int64)
        h = (domain[1::2] - domain[0::2]) / level_cells
        nwin_cells = phi - plo - 1
        wdom = domain.copy()
        wdom[4] = domain[4] + plo * h[2]
        wdom[5] = domain[4] + (phi - 1) * h[2]
        Mw = getRegularMesh(wdom, [level_cells[0], level_cells[1], nwin_cells])
        if kappa2 is None:
            shift = rel_shift * float(np.sum(4.0 / h ** 2))
            A = nodal_stencil_matrix(Mw, None, shift)
        else:
            A = nodal_stencil_matrix(Mw, None, 0.0).astype(np.complex128)
            A = sp.csc_matrix(A + sp.identity(A.shape[0], dtype=np.complex128) * (-kappa2 * (1.0 - 1j * gamma)))
        # the z-edges cut by the window ends belong to the global operator's diagonal
        plane = int((level_cells[0] + 1) * (level_cells[1] + 1))
        fix = np.zeros(A.shape[0])
        if plo > 0:
            fix[:plane] += 1.0 / h[2] ** 2
        if phi < level_cells[2] + 1:
            fix[-plane:] += 1.0 / h[2] ** 2
        A = sp.csc_matrix(A + sp.diags(fix))
        A.sort_indices()
        return A
    return build


def _window(own_lo, own_hi, n_planes, margin, align):
    lo = max(0, ((own_lo - margin) // align) * align)
    hi = min(n_planes, -((-(own_hi + margin)) // align) * align + 1)
    return lo, hi


def setup_slab_hierarchy(window_operator, domain, n_cells, param, rank, world, replicate_below=200000,
                         gather=None, verbose=False, rediscretise=False):
    """Distributed twin of ``MGsetup`` (MGsetup.jl:7-138) for rank ``rank``.

    window_operator(level_cells, plane_lo, plane_hi) -> principal sub-matrix of the operator on the mesh with
    ``level_cells`` cells.  Galerkin path (default): only the fine level uses it, coarser operators are windowed
    products R A P.  rediscretise=True: the multilevelOperatorConstructor path (MGsetup.jl:28,105-106 with a
    parameter-free PDE, restrictParams = identity): EVERY level's operator is window_operator on that level's mesh.
    gather(obj) -> list of obj from all ranks (host all-gather; identity list for world == 1).
    Returns a DistHierarchy; ``param`` supplies levels / relaxType / relaxParam / VAL."""
    n_cells = np.asarray(n_cells, dtype=np.int64)
    assert len(n_cells) == 3, "slab partition is defined for 3-D grids"
    VAL = param.VAL
    rVAL = np.float64
    if gather is None:
        assert world == 1
        gather = lambda o: [o]
    # ---- level geometry (every dimension coarsens independently; stop when none can) -------
    cells = [n_cells.copy()]
    for l in range(param.levels - 1):
        nn = cells[-1] + 1
        nxt = np.array([(k + 1) // 2 - 1 if (k % 2 == 1 and k > 2) else k - 1 for k in nn], dtype=np.int64)
        if np.all(nxt == cells[-1]):
            break
        cells.append(nxt)
    L = len(cells)
    nrows = [int(np.prod(c + 1)) for c in cells]
    # first replicated level k (0-based index into levels); at least the coarsest is replicated
    k = L - 1
    for l in range(1, L):
        if nrows[l] < replicate_below or cells[l][2] // world < 2 or cells[l][2] % 2 == 1:
            k = l
            break
    if world == 1:
        k = min(k, L - 1)
    nd = k                      # number of distributed levels: 0..k-1
    if rediscretise:
        return _setup_slab_rediscretised(window_operator, domain, n_cells, param, rank, world, gather, verbose,
                                         cells, nrows, nd, L)
    # ---- owned planes per level and the fine window -------------------------------------------
    own = [slab_planes(int(cells[l][2]), world) for l in range(nd + 1)]
    # Fine window.  Only the plane at a cut end of a window holds inexact rows, on EVERY level: the fine window is a
    # principal sub-matrix (its edge rows miss the columns outside), and coarse plane k of a windowed Galerkin
    # product is exact as soon as the fine planes 2k-2 .. 2k+2 lie in the window and the rows of 2k-1 .. 2k+1 are
    # exact, i.e. from the second coarse plane on.  Ghost planes are needed as column indices only, not as rows.
    # So the window must hold, for every level l <= nd, the owned planes of that level plus ONE plane on either
    # side, expressed in fine planes (x 2^l); the slab boundaries of different levels need not coincide
    # (slab_planes divides the cells of each level separately).  tests/test_dist_gloo.py compares bit for bit.
    align = 2 ** nd
    n_planes = int(cells[0][2]) + 1
    need_lo = min((own[l][rank][0] - 1) * 2 ** l for l in range(nd + 1))
    need_hi = max(own[l][rank][1] * 2 ** l for l in range(nd + 1))          # last plane needed (inclusive)
    wlo = max(0, (need_lo // align) * align)
    whi = min(n_planes, -((-need_hi) // align) * align + 1)
    # ---- windowed Galerkin hierarchy -----------------------------------------------------------
    A = sp.csc_matrix(window_operator(cells[0], wlo, whi))
    if A.dtype != VAL:
        A = sp.csc_matrix(A, dtype=VAL)
    AT = _csc(A.conj().T) if np.iscomplexobj(A.data) else _csc(A.T)
    relaxParam = param.relaxParam
    out = DistHierarchy(rank=rank, world=world, n_cells=n_cells, levels=L)
    win = (wlo, whi)
    AT_k_owned = None
    for l in range(nd):
        nn = cells[l] + 1
        plane_f = int(nn[0] * nn[1])
        ncn = cells[l + 1] + 1
        plane_c = int(ncn[0] * ncn[1])
        P1, _ = get1DFWInterp(int(nn[0]), False)
        P2, _ = get1DFWInterp(int(nn[1]), False)
        P3g, _ = get1DFWInterp(int(nn[2]), False)                 # global 1-D interpolation in z
        cwin = (win[0] // 2, (win[1] - 1) // 2 + 1)               # coarse planes covered by the window
        P3 = sp.csc_matrix(P3g[win[0]:win[1], cwin[0]:cwin[1]])
        P = sp.kron(P3, sp.kron(P2, P1, format="csc"), format="csc")
        P.sort_indices()
        RT = _csc(P.copy(), dtype=rVAL)
        RT.data *= 0.5 ** 3
        PT = _csc(P.T, dtype=rVAL)
        d_win = getRelaxPrec(AT, param.relaxType, relaxParam if not isinstance(relaxParam, (list, tuple, np.ndarray))
                             else relaxParam[l], VAL)
        Ac_T = galerkin(PT, AT, RT)
        if Ac_T.dtype != VAL:
            Ac_T = _csc(Ac_T, dtype=VAL)
        # ---- extract the owned rows with global column indices ------------------------------
        olo, ohi = own[l][rank]
        c_olo, c_ohi = own[l + 1][rank]
        r0, r1 = (olo - win[0]) * plane_f, (ohi - win[0]) * plane_f
        cr0, cr1 = (c_olo - cwin[0]) * plane_c, (c_ohi - cwin[0]) * plane_c
        n_glob, nc_glob = nrows[l], nrows[l + 1]

        def lift(Mcsc, row_shift, n_rows_global):
            Mcsc = sp.csc_matrix(Mcsc)
            return sp.csc_matrix((Mcsc.data, Mcsc.indices.astype(np.int64) + row_shift, Mcsc.indptr),
                                 shape=(n_rows_global, Mcsc.shape[1]))
        AT_own = lift(AT[:, r0:r1], win[0] * plane_f, n_glob)
        PT_own = lift(PT[:, r0:r1], cwin[0] * plane_c, nc_glob)
        RT_own = lift(RT[:, cr0:cr1], win[0] * plane_f, n_glob)
        roff = np.array([o[0] * plane_f for o in own[l]] + [n_glob], dtype=np.int64)
        croff = np.array([o[0] * plane_c for o in own[l + 1]] + [nc_glob], dtype=np.int64)
        out.dist_levels.append(DistLevel(n_global=n_glob, row_offsets=roff, AT=AT_own, PT=PT_own, RT=RT_own,
                                         d=np.ascontiguousarray(d_win[r0:r1]), nc_global=nc_glob,
                                         coarse_row_offsets=croff))
        if verbose:
            print(f"[rank {rank}] level {l + 1}: window planes {win}, owned planes {(olo, ohi)}, "
                  f"rows {r1 - r0}, nnz {AT_own.nnz}")
        if l == nd - 1:
            AT_k_owned = lift(Ac_T[:, cr0:cr1], cwin[0] * plane_c, nc_glob)
        AT = Ac_T
        win = cwin
    # ---- replicated levels: assemble level k globally, continue the ordinary setup ----------------
    from .mgdef import getMGparam
    from .mgsetup import MGsetup
    if nd == 0:
        A_k_T = AT if world == 1 else None
        assert world == 1, "at least one distributed level is needed for world > 1"
    else:
        pieces = gather(AT_k_owned)
        A_k_T = _csc(sp.hstack(pieces, format="csc"))
    rep = getMGparam(VAL, np.int64, L - nd, param.numCores, param.maxOuterIter, param.relativeTol,
                     param.relaxType, relaxParam if not isinstance(relaxParam, (list, tuple, np.ndarray))
                     else list(relaxParam[nd:]), param.relaxPre, param.relaxPost, param.cycleType,
                     param.coarseSolveType)
    Mk = getRegularMesh(domain, cells[nd])
    MGsetup(A_k_T, Mk, rep, 1)
    out.replicated = rep
    out.nd = nd
    out.cells = cells
    return out


def _setup_slab_rediscretised(window_operator, domain, n_cells, param, rank, world, gather, verbose, cells, nrows, nd, L):
    """Rediscretised hierarchy, rank's rows only: level l's owned rows come from the operator on the window of owned
    planes plus one plane on either side (a principal sub-matrix whose owned rows are complete); P from the 1-D
    interpolations restricted to the owned fine planes, R to the owned coarse planes."""
    VAL = param.VAL
    rVAL = np.float64
    relaxParam = param.relaxParam
    own = [slab_planes(int(cells[l][2]), world) for l in range(nd + 1)]
    out = DistHierarchy(rank=rank, world=world, n_cells=n_cells, levels=L)

    def lift(Mcsc, row_shift, n_rows_global):
        Mcsc = sp.csc_matrix(Mcsc)
        return sp.csc_matrix((Mcsc.data, Mcsc.indices.astype(np.int64) + row_shift, Mcsc.indptr),
                             shape=(n_rows_global, Mcsc.shape[1]))
    for l in range(nd):
        nn = cells[l] + 1
        ncn = cells[l + 1] + 1
        plane_f, plane_c = int(nn[0] * nn[1]), int(ncn[0] * ncn[1])
        n_glob, nc_glob = nrows[l], nrows[l + 1]
        olo, ohi = own[l][rank]
        c_olo, c_ohi = own[l + 1][rank]
        # ---- A: window = owned planes + one on either side -----------------------------------------------------
        wlo, whi = max(0, olo - 1), min(int(nn[2]), ohi + 1)
        A = sp.csc_matrix(window_operator(cells[l], wlo, whi))
        if A.dtype != VAL:
            A = sp.csc_matrix(A, dtype=VAL)
        AT = _csc(A.conj().T) if np.iscomplexobj(A.data) else _csc(A.T)
        d_win = getRelaxPrec(AT, param.relaxType, relaxParam if not isinstance(relaxParam, (list, tuple, np.ndarray))
                             else relaxParam[l], VAL)
        r0, r1 = (olo - wlo) * plane_f, (ohi - wlo) * plane_f
        AT_own = lift(AT[:, r0:r1], wlo * plane_f, n_glob)
        # ---- P rows of the owned fine planes, R rows of the owned coarse planes ---------------------------------
        P1, _ = get1DFWInterp(int(nn[0]), False)
        P2, _ = get1DFWInterp(int(nn[1]), False)
        P3g = sp.csr_matrix(get1DFWInterp(int(nn[2]), False)[0])
        P12 = sp.kron(P2, P1, format="csc")
        # fine planes olo..ohi-1 interpolate from coarse planes olo//2 .. (ohi-1+1)//2
        cp_lo, cp_hi = olo // 2, min(int(ncn[2]), (ohi - 1 + 1) // 2 + 1)
        Pown = sp.kron(sp.csc_matrix(P3g[olo:ohi, cp_lo:cp_hi]), P12, format="csc")      # (owned fine rows) x (coarse window)
        PT_own = lift(_csc(Pown.T, dtype=rVAL), cp_lo * plane_c, nc_glob)
        # coarse planes c_olo..c_ohi-1 restrict from fine planes 2c-1 .. 2c+1
        fp_lo, fp_hi = max(0, 2 * c_olo - 1), min(int(nn[2]), 2 * (c_ohi - 1) + 2)
        Pcw = sp.kron(sp.csc_matrix(P3g[fp_lo:fp_hi, c_olo:c_ohi]), P12, format="csc")    # (fine window) x (owned coarse)
        RT_own = _csc(Pcw, dtype=rVAL)
        RT_own.data *= 0.5 ** 3
        RT_own = lift(RT_own, fp_lo * plane_f, n_glob)
        roff = np.array([o[0] * plane_f for o in own[l]] + [n_glob], dtype=np.int64)
        croff = np.array([o[0] * plane_c for o in own[l + 1]] + [nc_glob], dtype=np.int64)
        out.dist_levels.append(DistLevel(n_global=n_glob, row_offsets=roff, AT=AT_own, PT=PT_own, RT=RT_own,
                                         d=np.ascontiguousarray(d_win[r0:r1]), nc_global=nc_glob,
                                         coarse_row_offsets=croff))
        if verbose:
            print(f"[rank {rank}] level {l + 1} (rediscretised): owned planes {(olo, ohi)}, rows {r1 - r0}, nnz {AT_own.nnz}")
    # ---- replicated levels: the ordinary rediscretised setup from level nd on -------------------------------------
    from .mgdef import getMGparam, getMultilevelOperatorConstructor
    from .mgsetup import MGsetup
    rep = getMGparam(VAL, np.int64, L - nd, param.numCores, param.maxOuterIter, param.relativeTol,
                     param.relaxType, relaxParam if not isinstance(relaxParam, (list, tuple, np.ndarray))
                     else list(relaxParam[nd:]), param.relaxPre, param.relaxPost, param.cycleType,
                     param.coarseSolveType)
    Mk = getRegularMesh(domain, cells[nd])

    def full_operator(mesh, _):
        c = np.asarray(mesh.n, dtype=np.int64)
        return window_operator(c, 0, int(c[2]) + 1)
    ctor = getMultilevelOperatorConstructor(None, full_operator, lambda mf, mc, pf, level: pf)
    MGsetup(ctor, Mk, rep, 1)
    out.replicated = rep
    out.nd = nd
    out.cells = cells
    return out
