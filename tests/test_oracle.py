"""CPU tests that pin the oracle as far as it can be pinned without a runnable reference:
every array kernel against an independent scipy/numpy evaluation, the operation counts of
SURVEY.md appendix A.2, the survey's sanity convergence values, and algebraic properties."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import make_problem
from oracle import cycle as oc
from oracle import kernels as K
from oracle import krylov


def _rand(shape, cplx, rng):
    a = rng.standard_normal(shape)
    if cplx:
        a = a + 1j * rng.standard_normal(shape)
    return np.asfortranarray(a)


@pytest.mark.parametrize("cplxA,cplxX", [(False, False), (True, True), (False, True)])
@pytest.mark.parametrize("nrhs", [1, 3])
def test_spmatmul_matches_scipy(cplxA, cplxX, nrhs):
    rng = np.random.default_rng(0)
    M = sp.random(40, 55, density=0.15, random_state=1, format="csc")
    if cplxA:
        M = M + 1j * sp.random(40, 55, density=0.15, random_state=2, format="csc")
    M = sp.csc_matrix(M)
    AT = K.CSCAdjoint(M)
    xs = (40,) if nrhs == 1 else (40, nrhs)
    ys = (55,) if nrhs == 1 else (55, nrhs)
    x = _rand(xs, cplxX, rng)
    y0 = _rand(ys, cplxX, rng)
    for alpha, beta in ((1.0, 0.0), (1.0, 1.0), (-1.0, 1.0), (-1.0, 0.0)):
        y = K.SpMatMul(alpha, AT, x, beta, y0.copy(order="F"), 2)
        ref = beta * y0 + alpha * (M.conj().T @ x)
        np.testing.assert_allclose(y, ref, rtol=1e-13, atol=1e-13)


def test_vector_kernels():
    rng = np.random.default_rng(1)
    for cplx in (False, True):
        x = _rand((1000,), cplx, rng)
        y = _rand((1000,), cplx, rng)
        d = _rand((1000,), cplx, rng)
        y2 = y.copy()
        K.addVectors(-0.75, x, y2)
        np.testing.assert_allclose(y2, y - 0.75 * x, rtol=1e-14)
        X = _rand((1000, 3), cplx, rng)
        R = _rand((1000, 3), cplx, rng)
        X2 = X.copy(order="F")
        K.scaleAdd(d, R, X2)
        np.testing.assert_allclose(X2, X + d[:, None] * R, rtol=1e-14)
        assert abs(K.norm(X) - np.linalg.norm(X)) < 1e-12 * np.linalg.norm(X)
        assert abs(K.dot(x, y) - np.vdot(x, y)) < 1e-11 * abs(np.vdot(x, y))


@pytest.mark.parametrize("cycle,pre,post", [('V', 2, 2), ('V', 1, 3), ('W', 1, 1), ('F', 2, 1)])
def test_operation_counts(cycle, pre, post):
    """with x = 0 on entry a level does nu1+nu2 passes over A, one over R,
    one over P per visit; visits per level: V 1, W 2^(l-1), F l."""
    A, AT, M, p, b = make_problem("poisson", [32, 32], 4, pre=pre, post=post, cycle=cycle, maxit=1)
    o = oc.OracleMG(p)
    MMG = oc.getMultigridPreconditioner(o, b)
    MMG(b)
    L = p.levels
    visits = {'V': [1] * (L - 1), 'W': [2 ** l for l in range(L - 1)], 'F': [l + 1 for l in range(L - 1)]}[cycle]
    # second calls of a W / F parent start from x != 0: one extra A pass each (MGcycle.jl:29-31).
    # W: half of the visits of every level >= 2; F: one per level >= 2 (only the F-type parent recurses twice)
    extra = {'V': 0, 'W': sum(v // 2 for v in visits[1:]), 'F': len(visits) - 1}[cycle]
    assert o.counters["A"] == sum(visits) * (pre + post) + extra
    assert o.counters["R"] == sum(visits) and o.counters["P"] == sum(visits)


def test_survey_sanity_values_2d():
    """G'G + 1e-4||.||_1 I, Jacobi 0.8, V(2,2), Galerkin, 128^2 cells, 4 levels,
    b = A*U(0,1)^n/||.|| with numpy default_rng(0)."""
    A, AT, M, p, b = make_problem("poisson", [128, 128], 4, maxit=6)
    o = oc.OracleMG(p)
    _, it, res = oc.solveMG(o, b, np.zeros_like(b))
    expect = [4.06e-2, 3.88e-3, 4.17e-4, 4.73e-5, 5.52e-6, 6.57e-7]
    np.testing.assert_allclose(res[1:] / res[0], expect, rtol=6e-3)


def test_survey_sanity_values_3d():
    A, AT, M, p, b = make_problem("poisson", [64, 64, 64], 4, maxit=6)
    o = oc.OracleMG(p)
    _, it, res = oc.solveMG(o, b, np.zeros_like(b))
    expect = [2.90e-2, 3.60e-3, 6.66e-4, 1.47e-4, 3.49e-5, 8.65e-6]
    np.testing.assert_allclose(res[1:] / res[0], expect, rtol=6e-3)


def test_cycle_is_linear_and_symmetric():
    """One V(nu,nu) cycle from x=0 is a linear operator B; for symmetric A, R = c P' and equal
    pre/post Jacobi sweeps it is symmetric."""
    A, AT, M, p, b = make_problem("poisson", [16, 16], 3, pre=2, post=2)
    o = oc.OracleMG(p)
    MMG = oc.getMultigridPreconditioner(o, b)
    rng = np.random.default_rng(2)
    u, v = rng.standard_normal(b.shape), rng.standard_normal(b.shape)
    Bu, Bv = MMG(u).copy(), MMG(v).copy()
    Buv = MMG(2.0 * u - 3.0 * v).copy()
    np.testing.assert_allclose(Buv, 2.0 * Bu - 3.0 * Bv, rtol=1e-10, atol=1e-12)
    assert abs(np.dot(Bu, v) - np.dot(u, Bv)) < 1e-10 * abs(np.dot(Bu, v))


def test_coarsest_only_hierarchy_is_exact():
    A, AT, M, p, b = make_problem("poisson", [8, 8], 1, maxit=1)
    o = oc.OracleMG(p)
    x, it, res = oc.solveMG(o, b, np.zeros_like(b))
    assert res[1] < 1e-12 * res[0]


def test_julia_pinv_rank_rule():
    H = np.diag([1.0, 1e-3, 0.0])
    P = oc.julia_pinv(H)
    np.testing.assert_allclose(np.diag(P), [1.0, 1e3, 0.0])
    H = np.diag([1.0, 1e-17])      # below eps*min(size)*max(S): dropped
    np.testing.assert_allclose(np.diag(oc.julia_pinv(H)), [1.0, 0.0])


def test_fgmres_relaxation_minimises_residual():
    """FGMRES.jl:48-126: the update minimises ||r0 - A Z t|| over t."""
    A, AT, M, p, b = make_problem("poisson", [16, 16], 2)
    o = oc.OracleMG(p)
    ATc = o.As[0]
    d = o.relaxPrecs[0]
    Afun = oc.getAfun(ATc, np.zeros_like(b), 2)
    y = np.zeros_like(b)

    def MM(v):
        y[...] = d * v
        return y
    x0 = np.zeros_like(b)
    x, rn = oc.FGMRES_relaxation(Afun, b.copy(), x0, 3, MM, 1e-30, 2)
    r = b - A @ x
    assert abs(np.linalg.norm(r) - rn[-1]) < 1e-8
    assert np.all(np.diff(rn) <= 1e-14)
    # not worse than three damped Jacobi steps with any single damping
    assert rn[-1] <= np.linalg.norm(b - A @ (d * b)) + 1e-14


def test_cg_matches_textbook_loop():
    A, AT, M, p, b = make_problem("poisson", [12, 12], 2)
    Ad = A.toarray()
    Afun = lambda v: Ad @ v
    x, flag, rel, it, resvec = krylov.cg(Afun, b, tol=1e-10, maxIter=400, x=np.zeros_like(b))
    assert flag == 0 and np.linalg.norm(b - Ad @ x) <= 1.01e-10 * np.linalg.norm(b)
    assert resvec[-1] == rel and len(resvec) == it


def test_fgmres_converges_and_counts_restarts():
    A, AT, M, p, b = make_problem("helmholtz", [10, 10], 2)
    Ad = A.toarray()
    Afun = lambda v: Ad @ v
    x, flag, rel, it, resvec = krylov.fgmres(Afun, b, 20, tol=1e-9, maxIter=30, x=np.zeros_like(b), flexible=True)
    assert flag == 0 and np.linalg.norm(b - Ad @ x) <= 1.5e-9 * np.linalg.norm(b)
    assert np.all(np.diff(resvec[:20]) <= 1e-12)  # monotone inside a restart


def test_blockcg_matches_columnwise_solution():
    A, AT, M, p, b = make_problem("poisson", [12, 12], 2, nrhs=4)
    Ad = A.toarray()
    Afun = lambda V: np.asfortranarray(Ad @ V)
    X, flag, rel, it, resmat = krylov.blockCG(Afun, b, X=np.zeros_like(b), tol=1e-10, maxIter=200)
    assert flag == 0
    np.testing.assert_allclose(Ad @ X, b, atol=2e-10)


def test_oracle_bicgstb_solves_nonsymmetric_system():
    """KrylovMethods.bicgstb restatement: converges on a nonsymmetric diagonally dominant system, with a
    Jacobi M1 and without, and the returned history is the true relative residual."""
    import scipy.sparse as sp
    from oracle import krylov
    rng = np.random.default_rng(5)
    n = 400
    A = sp.diags([-1.0 - 0.3 * rng.random(n - 1), 4.0 + rng.random(n), -1.0 + 0.3 * rng.random(n - 1)], [-1, 0, 1],
                 format="csr")
    b = rng.random(n)
    d = 1.0 / A.diagonal()
    for M1 in (None, lambda v: d * v):
        x, flag, rel, it, res = krylov.bicgstb(lambda v: A @ v, b, tol=1e-10, maxIter=200, M1=M1)
        assert flag in (0, -3) and it < 60
        assert np.linalg.norm(b - A @ x) / np.linalg.norm(b) <= 2e-10
        assert abs(res[-1] - np.linalg.norm(b - A @ x) / np.linalg.norm(b)) <= 1e-9
    # b = 0 and an exact initial guess
    assert krylov.bicgstb(lambda v: A @ v, np.zeros(n))[1] == -9
    xs = np.linalg.solve(A.toarray(), b)
    assert krylov.bicgstb(lambda v: A @ v, b, tol=1e-8, x=xs)[3] == 0


def test_block_krylov_with_one_column_equals_the_scalar_drivers():
    """blockFGMRES / blockBiCGSTB restatements (KrylovMethods, un-vendored): with one right-hand side they are the
    scalar fgmres / bicgstb - same iteration counts, flags and residual histories."""
    from oracle import krylov
    for kind in ("poisson", "helmholtz"):
        A, AT, M, p, b = make_problem(kind, [14, 14], 2)
        Ad = A.toarray()
        d = 1.0 / np.diag(Ad)
        Afun = lambda v: np.asfortranarray(Ad @ v)
        Mfun = lambda v: np.asfortranarray((d * v.T).T.copy())
        B = np.asfortranarray(b.reshape(-1, 1))
        x1, f1, r1, i1, h1 = krylov.fgmres(Afun, b, 80, tol=1e-8, maxIter=10, M=Mfun, x=np.zeros_like(b), flexible=True)
        X2, f2, r2, i2, h2 = krylov.blockFGMRES(Afun, B, 80, tol=1e-8, maxIter=10, M=Mfun, X=np.zeros_like(B),
                                                flexible=True)
        assert (f1, i1, len(h1)) == (f2, i2, len(h2)) and f1 == 0
        np.testing.assert_allclose(h2, h1, rtol=1e-6)
        np.testing.assert_allclose(X2[:, 0], x1, rtol=0, atol=1e-9 * np.linalg.norm(x1))
        # BiCGStab's short recurrences amplify rounding on the nearly singular operator: compare on a
        # well-conditioned one (the scalar code divides rho/rho1, the block code solves with R~'V)
        Aw = Ad + 0.5 * np.diag(np.diag(Ad))
        Awfun = lambda v: np.asfortranarray(Aw @ v)
        x1, f1, r1, i1, h1 = krylov.bicgstb(Awfun, b, tol=1e-9, maxIter=80, M1=Mfun, x=np.zeros_like(b))
        X2, f2, r2, i2, h2 = krylov.blockBiCGSTB(Awfun, B, tol=1e-9, maxIter=80, M1=Mfun, x=np.zeros_like(B))
        assert (f1, i1, len(h1)) == (f2, i2, len(h2)) and f1 in (0, -3) and i1 < 30
        np.testing.assert_allclose(h2, h1, rtol=1e-5)


@pytest.mark.parametrize("kind,nrhs", [("poisson", 4), ("helmholtz", 3)])
def test_block_krylov_solves_all_columns(kind, nrhs):
    """blockFGMRES (flexible and not) and blockBiCGSTB reach the Frobenius tolerance and report the true residual."""
    from oracle import krylov
    A, AT, M, p, b = make_problem(kind, [12, 12], 2, nrhs=nrhs)
    Ad = A.toarray()
    d = 1.0 / np.diag(Ad)
    Afun = lambda V: np.asfortranarray(Ad @ V)
    Mfun = lambda V: np.asfortranarray((d * V.T).T.copy())
    nb = np.linalg.norm(b)
    for flexible in (True, False):
        X, flag, rel, it, res = krylov.blockFGMRES(Afun, b, 40, tol=1e-9, maxIter=20, M=Mfun, X=np.zeros_like(b),
                                                   flexible=flexible)
        assert flag == 0 and np.linalg.norm(b - Ad @ X) <= 1.5e-9 * nb
        assert abs(res[-1] - np.linalg.norm(b - Ad @ X) / nb) <= 1e-10
    Aw = Ad + 0.5 * np.diag(np.diag(Ad))        # BiCGStab with a Jacobi M1 stalls on the nearly singular operator
    X, flag, rel, it, res = krylov.blockBiCGSTB(lambda V: np.asfortranarray(Aw @ V), b, tol=1e-9, maxIter=100,
                                                M1=Mfun, x=np.zeros_like(b))
    assert flag in (0, -3) and np.linalg.norm(b - Aw @ X) <= 1.5e-9 * nb
    assert abs(res[-1] - np.linalg.norm(b - Aw @ X) / nb) <= 1e-10


@pytest.mark.parametrize("kind,n,levels,cycle,relax", [("poisson", [32, 32], 3, 'V', "Jac"), ("poisson", [12, 12, 12], 3, 'W', "Jac"),
                                                       ("diffusion", [24, 24], 3, 'F', "SPAI"), ("helmholtz", [24, 24], 3, 'V', "Jac"),
                                                       ("helmholtz", [10, 10, 10], 3, 'W', "SPAI")])
def test_two_independent_restatements_agree(kind, n, levels, cycle, relax):
    """oracle/cycle.py (control flow + the C kernels of mg_kernels.c) and oracle/cycle32.py (plain scipy / numpy passes,
    written separately) restate the same lines of MGcycle.jl / SolveFuncs.jl: in double precision their per-cycle
    residual norms agree to rounding (different accumulation code, same operation order)."""
    from conftest import make_problem
    from oracle import cycle as oc, cycle32 as o32
    A, AT, M, p, b = make_problem(kind, n, levels, cycle=cycle, relax=relax, omega=0.8 if relax == "Jac" else 1.0, maxit=5)
    _, it_c, res_c = oc.solveMG(oc.OracleMG(p), b, np.zeros_like(b))
    _, it_n, res_n = o32.solveMG(o32.OracleMG32(p), b, np.zeros_like(b))
    assert it_c == it_n
    np.testing.assert_allclose(res_n, res_c, rtol=1e-11)
