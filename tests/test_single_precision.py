"""Single-precision hierarchies (getMGparam(Float32 / ComplexF32, ...), MGdef.jl:119,151, MGsetup.jl:31-33,79-82,
108-110) and the mixed-precision preconditioner under a double-precision Krylov method (SolveFuncs.jl:52-60).

Tolerances: a single-precision cycle carries a rounding error of a few eps_single = 6e-8 per pass; the per-cycle
residual norms of the device and of the CPU restatement (oracle/cycle32.py) must agree within 5e-4 relative while the
residual is above the single-precision floor, and the mixed-precision Krylov drivers must take the same number of
iterations with residual histories within 1e-2 relative."""
import numpy as np
import pytest

from conftest import make_problem

RTOL32 = 5e-4


def make_single(kind, n, levels, cycle='V', nrhs=1, maxit=5, relax="Jac", omega=0.8, tol=0.0):
    """The double-precision problem of conftest.make_problem and a Float32 / ComplexF32 hierarchy of the same
    matrix (the reference converts As[1] and every Galerkin product to VAL when `singlePrecision`)."""
    import multigrid_jl_b200 as mg
    A, AT, M, p64, b = make_problem(kind, n, levels, cycle=cycle, nrhs=nrhs, relax=relax, omega=omega)
    VAL = np.complex64 if p64.VAL == np.complex128 else np.float32
    p = mg.getMGparam(VAL, np.int64, levels, 8, maxit, tol, relax, omega, 2, 2, cycle)
    mg.MGsetup(AT, M, p, nrhs)
    return A, AT, M, p, b


# ---- CPU: host setup and the restatement itself -------------------------------------------------------------------
@pytest.mark.parametrize("kind,n", [("poisson", [32, 32]), ("helmholtz", [24, 24])])
def test_single_precision_setup_types(kind, n):
    A, AT, M, p, b = make_single(kind, n, 3)
    rT = np.float32
    assert p.singlePrecision
    assert all(a.dtype == p.VAL for a in p.As)
    assert all(m.dtype == rT for m in p.Ps + p.Rs)
    assert all(d.dtype == p.VAL for d in p.relaxPrecs)


def test_sa_amg_setup_single_precision_types():
    """SA_AMGsetup with a Float32 param: Ps / Rs are stored as real(VAL) (typed arrays, SA-AMG.jl:9-10), the Galerkin
    products are formed from those rounded operators, the aggregates are those of the double-precision setup."""
    import multigrid_jl_b200 as mg
    M = mg.getRegularMesh([0, 1, 0, 1], [24, 24])
    rng = np.random.default_rng(0)
    w = mg.edge_weights_from_cells(M, np.exp(rng.standard_normal(24 * 24)))
    A = mg.nodal_stencil_matrix(M, w, 1e-3)
    p32 = mg.getMGparam(np.float32, np.int64, 3, 8, 5, 1e-6, "SPAI", 1.0, 1, 1, 'W')
    p64 = mg.getMGparam(np.float64, np.int64, 3, 8, 5, 1e-6, "SPAI", 1.0, 1, 1, 'W')
    mg.SA_AMGsetup(A, p32, True, 1)
    mg.SA_AMGsetup(A, p64, True, 1)
    assert all(m.dtype == np.float32 for m in p32.As + p32.Ps + p32.Rs)
    assert all(d.dtype == np.float32 for d in p32.relaxPrecs)
    assert [a.shape for a in p32.As] == [a.shape for a in p64.As]
    for a32, a64 in zip(p32.aggregates, p64.aggregates):
        assert np.array_equal(a32, a64)
    for P32, P64 in zip(p32.Ps, p64.Ps):
        assert np.array_equal(P32.indices, P64.indices)
        np.testing.assert_allclose(P32.data, P64.data, rtol=1e-4, atol=1e-5)   # entries are O(1), some by cancellation
    import scipy.sparse as sp
    G = sp.csc_matrix((p32.Ps[0] @ p32.As[0]) @ p32.Rs[0])
    G.sort_indices()
    assert G.dtype == np.float32 and np.array_equal(G.data, p32.As[1].data)


def test_oracle32_cycle_converges_to_single_precision_floor():
    from oracle import cycle as oc, cycle32 as o32
    import multigrid_jl_b200 as mg
    A, AT, M, p, b = make_single("poisson", [32, 32], 3, maxit=8)
    o = o32.OracleMG32(p)
    x, it, res = o32.solveMG(o, b.astype(np.float32), np.zeros(b.shape, np.float32))
    assert x.dtype == np.float32 and it == 8
    # the first cycles contract like the double-precision cycle of the same matrix
    _, _, _, p64, _ = make_problem("poisson", [32, 32], 3, maxit=8)
    _, _, res64 = oc.solveMG(oc.OracleMG(p64), b, np.zeros_like(b))
    np.testing.assert_allclose(res[:3] / res[0], res64[:3] / res64[0], rtol=1e-3)
    assert res[-1] / res[0] < 1e-4            # and stagnates near the single-precision floor, not before


def test_oracle32_mixed_precision_pcg_matches_double():
    """A single-precision cycle is as good a preconditioner as the double-precision one (SolveFuncs.jl:52-60)."""
    from oracle import cycle as oc, cycle32 as o32, krylov, kernels as K
    A, AT, M, p, b = make_single("poisson", [48, 48], 3)
    _, _, _, p64, _ = make_problem("poisson", [48, 48], 3)
    Afun = oc.getAfun(K.CSCAdjoint(AT), np.zeros_like(b), 0)
    x, flag, _, it, resvec = krylov.cg(Afun, b, tol=1e-8, maxIter=30, M=o32.getMultigridPreconditioner(o32.OracleMG32(p)),
                                       x=np.zeros_like(b))
    o64 = oc.OracleMG(p64)
    x64, flag64, _, it64, _ = krylov.cg(Afun, b, tol=1e-8, maxIter=30, M=oc.getMultigridPreconditioner(o64, b),
                                        x=np.zeros_like(b))
    assert flag == 0 and flag64 == 0 and abs(it - it64) <= 1
    assert np.linalg.norm(b - A @ x) <= 2e-8 * np.linalg.norm(b)


# ---- GPU parity -----------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("kind,n,levels,cycle", [("poisson", [64, 64], 3, 'V'), ("poisson", [32, 32, 16], 3, 'W'),
                                                 ("diffusion", [24, 24, 24], 3, 'V'), ("helmholtz", [48, 48], 3, 'V'),
                                                 ("helmholtz", [16, 16, 16], 3, 'F')])
@pytest.mark.parametrize("tma", [False, True])
def test_solveMG_single_precision(kind, n, levels, cycle, tma):
    import multigrid_jl_b200 as mg
    from oracle import cycle32 as o32
    A, AT, M, p, b = make_single(kind, n, levels, cycle=cycle, maxit=4)
    bs = b.astype(p.VAL)
    o = o32.OracleMG32(p)
    x_ref, it_ref, res_ref = o32.solveMG(o, bs, np.zeros_like(bs))
    dev = mg.uploadHierarchy(p)
    if tma:
        dev.set_option("tma_min_rows", 0)     # ComplexF32 takes the TMA-staged dictionary kernel, Float32 the one-pass one
    x = np.zeros_like(bs)
    x, _, it = mg.solveMG(p, bs, x)
    assert x.dtype == p.VAL and it == it_ref
    np.testing.assert_allclose(p.last_resvec, res_ref, rtol=RTOL32)
    assert np.linalg.norm(x - x_ref) <= 1e-4 * np.linalg.norm(x_ref)


@pytest.mark.gpu
def test_single_precision_block():
    import multigrid_jl_b200 as mg
    from oracle import cycle32 as o32
    A, AT, M, p, b = make_single("poisson", [32, 32], 3, nrhs=4, maxit=3)
    bs = np.asfortranarray(b.astype(np.float32))
    o = o32.OracleMG32(p)
    x_ref, it_ref, res_ref = o32.solveMG(o, bs, np.zeros_like(bs))
    x = np.zeros_like(bs)
    x, _, it = mg.solveMG(p, bs, x)
    np.testing.assert_allclose(p.last_resvec, res_ref, rtol=RTOL32)


@pytest.mark.gpu
@pytest.mark.parametrize("kind,n,levels", [("poisson", [64, 64], 3), ("poisson", [24, 24, 24], 3), ("diffusion", [32, 32], 3)])
def test_mixed_precision_solveCG(kind, n, levels):
    """solveCG_MG(AT::Float64, param::MGparam{Float32}, b::Float64, x0::Float64)."""
    import multigrid_jl_b200 as mg
    from oracle import cycle as oc, cycle32 as o32, krylov, kernels as K
    A, AT, M, p, b = make_single(kind, n, levels, maxit=30, tol=1e-8)
    Afun = oc.getAfun(K.CSCAdjoint(AT), np.zeros_like(b), 0)
    x_ref, flag_ref, _, it_ref, res_ref = krylov.cg(Afun, b, tol=1e-8, maxIter=30,
                                                    M=o32.getMultigridPreconditioner(o32.OracleMG32(p)), x=np.zeros_like(b))
    x = np.zeros_like(b)
    x, _, it = mg.solveCG_MG(AT, p, b, x)
    assert x.dtype == np.float64
    assert p.last_flag == flag_ref == 0 and it == it_ref
    np.testing.assert_allclose(p.last_resvec, res_ref, rtol=1e-2)
    assert np.linalg.norm(b - A @ x) <= 2e-8 * np.linalg.norm(b)   # double-precision accuracy from a single-precision cycle


@pytest.mark.gpu
def test_mixed_precision_fgmres_complex():
    """solveGMRES_MG with a ComplexF32 hierarchy under ComplexF64 vectors (cfg5's types)."""
    import multigrid_jl_b200 as mg
    from oracle import cycle as oc, cycle32 as o32, krylov, kernels as K
    A, AT, M, p, b = make_single("helmholtz", [48, 48], 3, maxit=10, tol=1e-8)
    Afun = oc.getAfun(K.CSCAdjoint(AT), np.zeros_like(b), 0)
    ref = krylov.fgmres(Afun, b, 5, tol=1e-8, maxIter=10, M=o32.getMultigridPreconditioner(o32.OracleMG32(p)),
                        x=np.zeros_like(b), flexible=True)
    x = np.zeros_like(b)
    x, _, it, res = mg.solveGMRES_MG(AT, p, b, x, True, 5)
    assert x.dtype == np.complex128 and p.last_flag == 0
    assert len(res) == len(ref[4])
    np.testing.assert_allclose(res, ref[4], rtol=1e-2)
    assert np.linalg.norm(b - A @ x) <= 2e-8 * np.linalg.norm(b)


@pytest.mark.gpu
def test_mixed_precision_blockCG_and_rebuild():
    """nrhs > 1 through the mixed handle, then a change of nrhs and a new setup (the outer handle is rebuilt)."""
    import multigrid_jl_b200 as mg
    A, AT, M, p, b = make_single("poisson", [32, 32], 3, nrhs=4, maxit=30, tol=1e-8)
    X = np.zeros_like(b)
    X, _, it = mg.solveCG_MG(AT, p, b, X)
    assert it <= 12 and np.linalg.norm(b - A @ X) <= 1e-7 * np.linalg.norm(b)
    b1 = np.ascontiguousarray(b[:, 0])
    x1 = np.zeros_like(b1)
    x1, _, it1 = mg.solveCG_MG(AT, p, b1, x1)
    assert np.linalg.norm(b1 - A @ x1) <= 1e-7 * np.linalg.norm(b1)
    mg.MGsetup(AT, M, p, 1)                   # invalidates both device handles
    x2 = np.zeros_like(b1)
    x2, _, it2 = mg.solveCG_MG(AT, p, b1, x2)
    assert it2 == it1 and np.allclose(x1, x2, rtol=0, atol=1e-12)


@pytest.mark.gpu
def test_mixed_handle_refuses_cycles():
    import multigrid_jl_b200 as mg
    from multigrid_jl_b200.device import DeviceHierarchy, MGB200Error
    A, AT, M, p, b = make_single("poisson", [32, 32], 3)
    dev = mg.uploadHierarchy(p)
    outer = DeviceHierarchy.mixed_over(dev, AT)
    with pytest.raises(MGB200Error):
        outer.solveMG(b, np.zeros_like(b), 1e-6, 3)
    outer.destroy()
