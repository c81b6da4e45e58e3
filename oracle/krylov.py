"""TEST INFRASTRUCTURE - restatement of the Krylov drivers the reference calls.

The drivers live in KrylovMethods.jl 0.6.0 (Manifest.toml:35-41, git-tree-sha1
ceb12d552f1d2a1d6395c771cee0e0bc3b069e08), an un-vendored dependency that is NOT
under /root/reference.  The algorithms below restate the published package from
memory and are anchored on the reference's call sites
(src/Multigrid/SolveFuncs.jl:93-130, MGcycle.jl:164-174).  PARITY UNPINNED:
"same Krylov iteration count" is therefore defined against THIS restatement.

Calling convention kept from the package: ``A`` and ``M`` are functions that
may return aliased buffers (SolveFuncs.jl:59,66-69), so results are copied
where the package copies them.
"""
from __future__ import annotations

import numpy as np

from . import kernels as K


def cg(A, b, tol=1e-2, maxIter=100, M=None, x=None, out=0):
    """KrylovMethods.cg: preconditioned CG, stop on ||r||/||b|| <= tol.
    Returns (x, flag, relres, iter, resvec)."""
    n = b.shape[0]
    M = (lambda v: v.copy()) if M is None else M
    if K.norm(b) == 0:
        return np.zeros(n, dtype=b.dtype), -9, 0.0, 0, np.array([0.0])
    if x is None or x.size == 0:
        x = np.zeros(n, dtype=b.dtype)
        r = b.copy()
    else:
        r = b - A(x)
    z = M(r)
    p = z.copy()
    nr0 = K.norm(b)
    resvec = np.zeros(maxIter)
    flag = -1
    lastIter = 0
    for it in range(1, maxIter + 1):
        lastIter = it
        Ap = A(p)
        gamma = K.dot(r, z)
        alpha = gamma / K.dot(p, Ap)
        if alpha == np.inf or alpha < 0:
            flag = -2
            break
        K.addVectors(alpha, p, x)            # x += alpha*p
        K.addVectors(-alpha, Ap, r)          # r -= alpha*Ap
        resvec[it - 1] = K.norm(r) / nr0
        if resvec[it - 1] <= tol:
            flag = 0
            break
        z = M(r)
        beta = K.dot(z, r) / gamma
        p *= beta                            # p = z + beta*p  (scal! then axpy!)
        K.addVectors(1.0, z, p)
    return x, flag, resvec[lastIter - 1], lastIter, resvec[:lastIter].copy()


def bicgstb(A, b, tol=1e-6, maxIter=100, M1=None, M2=None, x=None, out=0):
    """KrylovMethods.bicgstb: van der Vorst's preconditioned BiCGStab in the form of the "Templates"
    book (p_hat = M2(M1(p)), v = A p_hat, s = r - alpha v, half-step exit on ||s||/||b|| < tol,
    s_hat = M2(M1(s)), t = A s_hat, omega = <t,s>/<t,t>).  Returns (x, flag, relres, iter, resvec)
    with resvec[0] the initial relative residual; flag -3 = converged at the half step, in which case
    ``iter`` counts the completed full iterations (the reference adds one preconditioner application
    for it: nprec = 2*iter + (flag == -3), SolveFuncs.jl:97)."""
    n = b.shape[0]
    M1 = (lambda v: v.copy()) if M1 is None else M1
    M2 = (lambda v: v) if M2 is None else M2
    if K.norm(b) == 0:
        return np.zeros(n, dtype=b.dtype), -9, 0.0, 0, np.array([0.0])
    if x is None or x.size == 0:
        x = np.zeros(n, dtype=b.dtype)
        r = b.copy()
    else:
        r = b - A(x)
    bnrm2 = K.norm(b)
    err = K.norm(r) / bnrm2
    resvec = np.zeros(maxIter + 1)
    resvec[0] = err
    if err < tol:
        return x, 0, err, 0, resvec[:1].copy()
    omega = alpha = rho1 = 1.0
    r_tld = r.copy()
    p = v = None
    flag = -1
    it = 0
    for it in range(1, maxIter + 1):
        rho = K.dot(r_tld, r)
        if rho == 0.0:
            flag = -2
            break
        if it > 1:
            beta = (rho / rho1) * (alpha / omega)
            p = r + beta * (p - omega * v)
        else:
            p = r.copy()
        p_hat = np.array(M2(M1(p)), copy=True)
        v = np.array(A(p_hat), copy=True)
        alpha = rho / K.dot(r_tld, v)
        s = r - alpha * v
        snorm = K.norm(s) / bnrm2
        if snorm < tol:
            x = x + alpha * p_hat
            resvec[it] = snorm
            return x, -3, snorm, it - 1, resvec[:it + 1].copy()
        s_hat = M2(M1(s))
        t = np.array(A(s_hat), copy=True)
        omega = K.dot(t, s) / K.dot(t, t)
        x = x + (alpha * p_hat + omega * s_hat)
        r = s - omega * t
        err = K.norm(r) / bnrm2
        resvec[it] = err
        if err <= tol:
            flag = 0
            break
        if omega == 0.0:
            flag = -2
            break
        rho1 = rho
    return x, flag, resvec[it], it, resvec[:it + 1].copy()


def _colnorms(X):
    return np.sqrt(np.sum((X.conj() * X).real, axis=0))


def blockCG(A, B, X=None, M=None, maxIter=20, tol=1e-2, ortho=False, pinvTol=None, out=0):
    """KrylovMethods.blockCG (O'Leary block CG): stop when the MAXIMUM column
    relative residual is <= tol.  Returns (X, flag, relres, iter, resmat)."""
    n, nrhs = B.shape
    M = (lambda v: v.copy()) if M is None else M
    if pinvTol is None:
        pinvTol = np.finfo(np.float64).eps * n
    if K.norm(np.asfortranarray(B)) == 0:
        return np.zeros_like(B), -9, 0.0, 0, np.array([0.0])
    if X is None:
        X = np.zeros((n, nrhs), dtype=B.dtype, order="F")
    R = np.array(B, order="F", copy=True)
    if not np.all(X == 0):
        R -= A(X)
    Z = M(R)
    P = np.array(Z, order="F", copy=True)
    nB = _colnorms(B)
    resmat = np.zeros((maxIter, nrhs))
    flag = -1
    it = 0
    for it in range(1, maxIter + 1):
        Q = A(P)
        PTQ = P.conj().T @ Q
        pinvPTQ = np.linalg.pinv(PTQ, rcond=pinvTol)
        Alpha = pinvPTQ @ (P.conj().T @ R)
        X += P @ Alpha
        R -= Q @ Alpha
        resmat[it - 1, :] = _colnorms(R) / nB
        if resmat[it - 1, :].max() <= tol:
            flag = 0
            break
        Z = M(R)
        Beta = -pinvPTQ @ (Q.conj().T @ Z)
        P = np.asfortranarray(Z + P @ Beta)
    return X, flag, resmat[it - 1, :].max(), it, resmat[:it, :].copy()


def fgmres(A, b, restrt, tol=1e-2, maxIter=100, M=None, x=None, out=0, flexible=False):
    """KrylovMethods.fgmres: restarted, right-preconditioned (F)GMRES.
    Orthogonalisation is classical Gram-Schmidt as two gemv calls (t = V'w; w -= V t);
    the small least-squares problem is solved every inner step for the residual
    estimate; ``maxIter`` counts restarts; stop on ||r||/||b|| <= tol.
    Returns (x, flag, relres, iter, resvec)."""
    n = b.shape[0]
    T = b.dtype
    M = (lambda v: v.copy()) if M is None else M
    if K.norm(b) == 0.0:
        return np.zeros(n, dtype=T), -9, 0.0, 0, np.array([0.0])
    if x is None or x.size == 0:
        x = np.zeros(n, dtype=T)
        r = b.copy()
    else:
        r = b.copy()
        r -= A(x)
    rnorm0 = K.norm(b)
    err = K.norm(r) / rnorm0
    if err < tol:
        return x, 0, err, 0, np.array([err])
    restrt = min(restrt, n - 1)
    V = np.zeros((n, restrt), dtype=T, order="F")
    Z = np.zeros((n, restrt), dtype=T, order="F") if flexible else None
    resvec = np.zeros(restrt * maxIter)
    flag = -1
    counter = 0
    it = 0
    while it < maxIter:
        it += 1
        H = np.zeros((restrt + 1, restrt), dtype=T)
        xi = np.zeros(restrt + 1, dtype=T)
        V[...] = 0
        if flexible:
            Z[...] = 0
        betta = K.norm(r)
        xi[0] = betta
        w = r * (1.0 / betta)
        V[:, 0] = w
        for j in range(restrt):
            z = M(w)
            if flexible:
                Z[:, j] = z
            w = A(z)
            counter += 1
            t = V.conj().T @ w                       # gemv 'C'
            H[:restrt, j] = t
            w = w - V @ t                            # gemv 'N'
            betta = K.norm(np.ascontiguousarray(w))
            H[j + 1, j] = betta
            w = w * (1.0 / betta)
            if j + 1 < restrt:
                V[:, j + 1] = w
            Hj = H[:j + 2, :j + 1]
            y = np.linalg.lstsq(Hj, xi[:j + 2], rcond=None)[0]
            err = np.linalg.norm(Hj @ y - xi[:j + 2]) / rnorm0
            resvec[counter - 1] = err
            if err <= tol:
                flag = 0
                break
        y = np.linalg.pinv(H) @ xi
        if flexible:
            w = Z @ y
        else:
            w = M(V @ y).copy()
        x = x + w
        if flag == 0:
            break
        if it < maxIter:
            r = b.copy()
            r -= A(x)
    return x, flag, resvec[counter - 1], it, resvec[:counter].copy()


def blockFGMRES(A, B, restrt, tol=1e-2, maxIter=100, M=None, X=None, out=0, flexible=False):
    """KrylovMethods.blockFGMRES: the block analogue of ``fgmres`` above - block Arnoldi with one n x m block per
    inner step, classical block Gram-Schmidt against the whole basis as two gemm calls (T = V'W; W -= V T),
    a Householder QR of the new block (its R factor is the sub-diagonal block of H), the small least-squares
    problem min || H Y - Xi ||_F solved every inner step for the residual estimate, Frobenius norms, stop on
    ||R||_F / ||B||_F <= tol; ``maxIter`` counts restarts.  Returns (X, flag, relres, iter, resvec)."""
    n, m = B.shape
    T = B.dtype
    M = (lambda v: v.copy()) if M is None else M
    if np.linalg.norm(B) == 0.0:
        return np.zeros((n, m), dtype=T, order="F"), -9, 0.0, 0, np.array([0.0])
    if X is None or X.size == 0:
        X = np.zeros((n, m), dtype=T, order="F")
        R = np.array(B, order="F", copy=True)
    else:
        X = np.array(X, order="F", copy=True)
        R = np.array(B, order="F", copy=True)
        R -= A(X)
    rnorm0 = np.linalg.norm(B)
    err = np.linalg.norm(R) / rnorm0
    if err < tol:
        return X, 0, err, 0, np.array([err])
    restrt = min(restrt, n - 1)
    Vbig = np.zeros((n, m * restrt), dtype=T, order="F")
    Zbig = np.zeros((n, m * restrt), dtype=T, order="F") if flexible else None
    resvec = np.zeros(restrt * maxIter)
    flag = -1
    counter = 0
    it = 0
    while it < maxIter:
        it += 1
        H = np.zeros(((restrt + 1) * m, restrt * m), dtype=T)
        xi = np.zeros(((restrt + 1) * m, m), dtype=T)
        Vbig[...] = 0
        if flexible:
            Zbig[...] = 0
        W, betta = np.linalg.qr(R)
        xi[:m, :] = betta
        jdone = 0
        for j in range(restrt):
            cs = slice(j * m, (j + 1) * m)
            Vbig[:, cs] = W
            Z = M(np.asfortranarray(W))
            if flexible:
                Zbig[:, cs] = Z
            W = np.array(A(np.asfortranarray(Z)), order="F", copy=True)
            counter += 1
            Tm = Vbig.conj().T @ W                       # gemm 'C','N'
            H[:restrt * m, cs] = Tm
            W = W - Vbig @ Tm                            # gemm 'N','N'
            W, betta = np.linalg.qr(W)
            H[(j + 1) * m:(j + 2) * m, cs] = betta
            Hj = H[:(j + 2) * m, :(j + 1) * m]
            y = np.linalg.lstsq(Hj, xi[:(j + 2) * m, :], rcond=None)[0]
            err = np.linalg.norm(Hj @ y - xi[:(j + 2) * m, :]) / rnorm0
            resvec[counter - 1] = err
            jdone = j + 1
            if err <= tol:
                flag = 0
                break
        y = np.linalg.lstsq(H[:(jdone + 1) * m, :jdone * m], xi[:(jdone + 1) * m, :], rcond=None)[0]
        if flexible:
            W = Zbig[:, :jdone * m] @ y
        else:
            W = np.array(M(np.asfortranarray(Vbig[:, :jdone * m] @ y)), copy=True)
        X = np.asfortranarray(X + W)
        if flag == 0:
            break
        if it < maxIter:
            R = np.array(B, order="F", copy=True)
            R -= A(X)
    return X, flag, resvec[counter - 1], it, resvec[:counter].copy()


def blockBiCGSTB(A, b, tol=1e-6, maxIter=100, M1=None, M2=None, x=None, out=0):
    """KrylovMethods.blockBiCGSTB: block BiCGStab for several right-hand sides (El Guennouni, Jbilou and Sadok,
    ETNA 16, 2003) written like ``bicgstb`` above: P_hat = M2(M1(P)), V = A P_hat, the m x m system
    (R~'V) alpha = R~'R, S = R - V alpha, half-step exit on ||S||_F/||B||_F < tol (flag -3), S_hat = M2(M1(S)),
    T = A S_hat, omega = <T,S>_F/<T,T>_F, X += P_hat alpha + omega S_hat, R = S - omega T,
    (R~'V) beta = -R~'T, P = R + (P - omega V) beta.  For m = 1 this is exactly ``bicgstb``.
    Returns (X, flag, relres, iter, resvec)."""
    n, m = b.shape
    T = b.dtype
    M1 = (lambda v: v.copy()) if M1 is None else M1
    M2 = (lambda v: v) if M2 is None else M2
    if np.linalg.norm(b) == 0:
        return np.zeros((n, m), dtype=T, order="F"), -9, 0.0, 0, np.array([0.0])
    if x is None or x.size == 0:
        x = np.zeros((n, m), dtype=T, order="F")
        r = np.array(b, order="F", copy=True)
    else:
        x = np.array(x, order="F", copy=True)
        r = np.array(b, order="F", copy=True)
        r -= A(x)
    bnrm2 = np.linalg.norm(b)
    err = np.linalg.norm(r) / bnrm2
    resvec = np.zeros(maxIter + 1)
    resvec[0] = err
    if err < tol:
        return x, 0, err, 0, resvec[:1].copy()
    r_tld = r.copy()
    p = r.copy()
    flag = -1
    it = 0
    for it in range(1, maxIter + 1):
        p_hat = np.array(M2(M1(np.asfortranarray(p))), order="F", copy=True)
        v = np.array(A(p_hat), order="F", copy=True)
        G = r_tld.conj().T @ v
        if np.linalg.matrix_rank(G) < m:
            flag = -2
            break
        alpha = np.linalg.solve(G, r_tld.conj().T @ r)
        s = np.asfortranarray(r - v @ alpha)
        snorm = np.linalg.norm(s) / bnrm2
        if snorm < tol:
            x = np.asfortranarray(x + p_hat @ alpha)
            resvec[it] = snorm
            return x, -3, snorm, it - 1, resvec[:it + 1].copy()
        s_hat = np.array(M2(M1(np.asfortranarray(s))), order="F", copy=True)
        t = np.array(A(s_hat), order="F", copy=True)
        omega = np.vdot(t, s) / np.vdot(t, t)
        x = np.asfortranarray(x + (p_hat @ alpha + omega * s_hat))
        r = np.asfortranarray(s - omega * t)
        err = np.linalg.norm(r) / bnrm2
        resvec[it] = err
        if err <= tol:
            flag = 0
            break
        if omega == 0.0:
            flag = -2
            break
        beta = np.linalg.solve(G, -(r_tld.conj().T @ t))
        p = np.asfortranarray(r + (p - omega * v) @ beta)
    return x, flag, resvec[it], it, resvec[:it + 1].copy()
