"""Regular-mesh stand-ins for the jInv.Mesh pieces the multigrid path reads.

The reference only touches ``Mesh.n`` (cells), ``Mesh.domain``, ``Mesh.h`` and
``Mesh.dim`` (reference: src/Multigrid/MGsetup.jl:35,60,96) and builds test
matrices with ``getNodalGradientMatrix`` (test/Multigrid/testGMGRAPforPoisson.jl:11).
jInv itself is an un-vendored dependency (Manifest.toml:149-155), so these are
our own definitions of the synthetic operators named in SURVEY.md section 8(d).

Ordering is lexicographic with x fastest: node (i1,i2,i3) -> i1 + nn1*(i2 + nn2*i3).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


class RegularMesh:
    """domain = [x0,x1,y0,y1(,z0,z1)], n = cells per dimension."""

    def __init__(self, domain, n):
        self.domain = np.asarray(domain, dtype=np.float64)
        self.n = np.asarray(n, dtype=np.int64).copy()
        self.dim = len(self.n)
        assert len(self.domain) == 2 * self.dim
        self.h = (self.domain[1::2] - self.domain[0::2]) / self.n

    def __repr__(self):
        return f"RegularMesh(domain={self.domain.tolist()}, n={self.n.tolist()})"


def getRegularMesh(domain, n):
    return RegularMesh(domain, n)


def _ddx(n):
    """n x (n+1) forward difference with rows [-1 1]."""
    return sp.diags([-np.ones(n), np.ones(n)], [0, 1], shape=(n, n + 1), format="csr")


def getNodalGradientMatrix(M: RegularMesh):
    """Stacked kron's of 1-D differences/h on the nodal grid (edges x nodes)."""
    n, h = M.n, M.h
    I = [sp.identity(k + 1, format="csr") for k in n]
    D = [_ddx(int(k)) / hk for k, hk in zip(n, h)]
    if M.dim == 2:
        blocks = [sp.kron(I[1], D[0]), sp.kron(D[1], I[0])]
    else:
        blocks = [sp.kron(I[2], sp.kron(I[1], D[0])),
                  sp.kron(I[2], sp.kron(D[1], I[0])),
                  sp.kron(D[2], sp.kron(I[1], I[0]))]
    return sp.vstack(blocks, format="csr")


def getNodalLaplacianMatrix(M: RegularMesh):
    G = getNodalGradientMatrix(M)
    return (G.T @ G).tocsc()


def edge_weights_from_cells(M: RegularMesh, sigma):
    """Edge weights for G' diag(w) G: arithmetic mean of the cell values of the
    cells that share the edge (our definition; jInv's own averaging is not
    available here).  Returns the per-direction weight arrays flattened in the
    row order of getNodalGradientMatrix."""
    n = [int(k) for k in M.n]
    s = np.asarray(sigma, dtype=np.float64).reshape(n[::-1])  # [z,]y,x
    out = []
    for d in range(M.dim):
        # edges along dimension d: n[d] edges in d, nodes (n+1) in the others
        acc = s
        cnt = np.ones_like(s)
        for e in range(M.dim):
            if e == d:
                continue
            ax = M.dim - 1 - e
            pad = [(0, 0)] * M.dim
            pad[ax] = (1, 1)
            a = np.pad(acc, pad)
            c = np.pad(cnt, pad)
            sl_lo = [slice(None)] * M.dim
            sl_hi = [slice(None)] * M.dim
            sl_lo[ax] = slice(0, -1)
            sl_hi[ax] = slice(1, None)
            acc = a[tuple(sl_lo)] + a[tuple(sl_hi)]
            cnt = c[tuple(sl_lo)] + c[tuple(sl_hi)]
        out.append((acc / cnt).ravel())
    return out


def getNodalDivSigGradMatrix(M: RegularMesh, sigma):
    """G' diag(w) G with w the edge average of the cell coefficient sigma."""
    G = getNodalGradientMatrix(M)
    w = np.concatenate(edge_weights_from_cells(M, sigma))
    return (G.T @ sp.diags(w) @ G).tocsc()


def nodal_stencil_matrix(M: RegularMesh, weights=None, diag_shift=0.0, dtype=np.float64):
    """Direct assembly of A = G' diag(w) G + diag_shift*I (7-/5-point nodal
    stencil, homogeneous Neumann) without forming G.  Used for the large
    configs where the sparse product G'G would dominate setup time.  Returns a
    CSC matrix with sorted indices; A is symmetric so these are also the CSR
    arrays."""
    n = [int(k) for k in M.n]
    nn = [k + 1 for k in n]
    dim = M.dim
    N = int(np.prod(nn))
    shape_rev = nn[::-1]
    if weights is None:
        weights = []
        for d in range(dim):
            e = list(nn)
            e[d] = n[d]
            weights.append(np.ones(int(np.prod(e))))
    diag = np.zeros(shape_rev, dtype=np.float64)
    lo, hi = [], []  # coupling to the lower / upper neighbour in each dim (0 at boundary)
    for d in range(dim):
        e = list(nn)
        e[d] = n[d]
        w = np.asarray(weights[d], dtype=np.float64).reshape(e[::-1]) / (M.h[d] ** 2)
        ax = dim - 1 - d
        pad = [(0, 0)] * dim
        pad[ax] = (1, 1)
        wp = np.pad(w, pad)
        sl_lo = [slice(None)] * dim
        sl_hi = [slice(None)] * dim
        sl_lo[ax] = slice(0, -1)
        sl_hi[ax] = slice(1, None)
        wl = wp[tuple(sl_lo)]  # weight of the edge below the node
        wh = wp[tuple(sl_hi)]  # weight of the edge above the node
        diag += wl + wh
        lo.append(wl.ravel())
        hi.append(wh.ravel())
    diag = diag.ravel() + diag_shift
    strides = [1]
    for d in range(dim - 1):
        strides.append(strides[-1] * nn[d])
    # ascending column order within a row: -s_{dim-1}, ..., -s_0, 0, +s_0, ..., +s_{dim-1}
    cols, vals, present = [], [], []
    idx = np.arange(N, dtype=np.int64)
    for d in reversed(range(dim)):
        cols.append(idx - strides[d]); vals.append(-lo[d]); present.append(lo[d] != 0)
    cols.append(idx); vals.append(diag); present.append(np.ones(N, dtype=bool))
    for d in range(dim):
        cols.append(idx + strides[d]); vals.append(-hi[d]); present.append(hi[d] != 0)
    present = np.stack(present, axis=1)
    cols = np.stack(cols, axis=1)[present]
    vals = np.stack(vals, axis=1)[present].astype(dtype)
    indptr = np.zeros(N + 1, dtype=np.int64)
    np.cumsum(present.sum(axis=1), out=indptr[1:])
    A = sp.csc_matrix((vals, cols, indptr), shape=(N, N))
    A.has_sorted_indices = True
    return A


def poisson_shifted(M: RegularMesh, rel_shift=1e-4):
    """A = G'G + rel_shift*||G'G||_1*I (cf. testGMGRAPforPoisson.jl:11-13);
    ||G'G||_1 = sum_d 4/h_d^2 for the Neumann nodal Laplacian."""
    norm1 = float(np.sum(4.0 / M.h ** 2))
    return nodal_stencil_matrix(M, None, rel_shift * norm1)


def helmholtz_shifted(M: RegularMesh, kappa2, gamma=0.5):
    """Complex shifted Laplacian A = G'G - kappa^2 (1 - gamma i) I."""
    A = nodal_stencil_matrix(M, None, 0.0).astype(np.complex128)
    A = A + sp.identity(A.shape[0], dtype=np.complex128, format="csc") * (-kappa2 * (1.0 - 1j * gamma))
    A.sort_indices()
    return A.tocsc()
