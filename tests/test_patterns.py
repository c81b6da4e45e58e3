"""Stencil dictionary (csrc/pattern.cuh): the host-side row deduplication must reproduce the CSR
arrays bit for bit (CPU), and the device kernels that use it must give results identical to the
CSR-stream kernels (GPU)."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import make_problem


def _rebuild(M, pat):
    """CSR arrays of the operator from the dictionary form."""
    M = sp.csc_matrix(M)
    M.sort_indices()
    n = M.shape[1]
    cols, vals, ptr = [], [], [0]
    for i in range(n):
        p = pat["pid"][i]
        k0, k1 = pat["pat_off"][p], pat["pat_off"][p + 1]
        base = i if pat["row_relative"] else pat["c0"][i]
        cols.extend(base + pat["delta"][k0:k1])
        vals.extend(pat["val"][k0:k1])
        ptr.append(len(cols))
    return np.array(ptr), np.array(cols), np.array(vals)


@pytest.mark.parametrize("n,levels", [([16, 16], 3), ([8, 8, 8], 3), ([12, 10, 6], 2)])
def test_host_patterns_reproduce_csr(n, levels):
    from multigrid_jl_b200 import device
    A, AT, M, p, b = _cpu_problem(n, levels)
    for l in range(levels - 1):
        for name, mat, rowrel in (("A", p.As[l], True), ("P", p.Ps[l], False), ("R", p.Rs[l], False)):
            mat = sp.csc_matrix(mat)
            mat.sort_indices()
            pat = device.host_build_patterns(mat)
            if pat is None:      # tiny coarse matrices are not worth a dictionary
                assert mat.shape[1] < 8 * 27
                continue
            assert pat["row_relative"] == rowrel, (name, l)
            ptr, cols, vals = _rebuild(mat, pat)
            assert np.array_equal(ptr, mat.indptr)
            assert np.array_equal(cols, mat.indices)
            assert np.array_equal(vals.view(np.int64), np.ascontiguousarray(mat.data, dtype=np.float64).view(np.int64))
            assert len(pat["pat_off"]) - 1 <= 27 if mat.shape[0] == mat.shape[1] else True


def test_host_patterns_reject_unstructured():
    from multigrid_jl_b200 import device
    rng = np.random.default_rng(3)
    M = sp.random(3000, 3000, density=0.003, random_state=rng, format="csc") + sp.identity(3000, format="csc")
    assert device.host_build_patterns(M) is None
    # structured sparsity but row-dependent values: no value dictionary either
    T = sp.diags([rng.random(2999), rng.random(3000) + 2.0, rng.random(2999)], [-1, 0, 1], format="csc")
    assert device.host_build_patterns(T) is None


def test_host_patterns_empty_rows_and_caps():
    from multigrid_jl_b200 import device
    n = 4000
    d = np.ones(n)
    d[::7] = 0.0
    M = sp.diags([d], [0], format="csc")
    M.eliminate_zeros()
    pat = device.host_build_patterns(M)
    assert pat is not None and len(pat["pat_off"]) - 1 == 2
    ptr, cols, vals = _rebuild(M, pat)
    M.sort_indices()
    assert np.array_equal(ptr, M.indptr) and np.array_equal(cols, M.indices)
    # a cap of one pattern cannot hold two
    assert device.host_build_patterns(M, max_patterns=1) is None


def _check_plan(plan, tile, al):
    """Every dictionary offset lies in exactly the window its stage offset points into; windows are 16-byte
    granular, disjoint in the stage buffer and long enough for a whole tile."""
    wins = plan["windows"]
    assert plan["total"] == sum(w[1] for w in wins)
    sb = 0
    for lo, ln, base in wins:
        assert lo % al == 0 and ln % al == 0 and base == sb
        sb += ln
    for d, so in zip(list(plan["delta"]) + [0], list(plan["soff"]) + [plan["centre"]]):
        hit = [(lo, ln, base) for lo, ln, base in wins if base <= so < base + ln]
        assert len(hit) == 1
        lo, ln, base = hit[0]
        assert so - base == d - lo                     # x[row + d] of thread t sits at stage[so + t]
        assert 0 <= so - base and (so - base) + tile <= ln   # ... for every t < tile


@pytest.mark.parametrize("elem_bytes,al", [(8, 2), (4, 4), (16, 2)])
def test_tma_plan_merges_windows_of_a_3d_stencil(elem_bytes, al, monkeypatch):
    """7-point and 27-point operators on a 41^3-node grid, tile 256: the y-neighbour lines (gap 41-2 < tile) share the
    window of the centre line, the z-planes (gap 41^2 - ... > tile) keep their own: 3 windows.  With the round-1
    threshold (gap 32) the same matrices need 5 and 9."""
    import multigrid_jl_b200 as mg
    from multigrid_jl_b200 import device
    monkeypatch.delenv("MGB200_TMA_GAP", raising=False)
    M = mg.getRegularMesh([0, 1, 0, 1, 0, 1], [40, 40, 40])
    A = mg.poisson_shifted(M, 1e-4)
    p = mg.getMGparam(np.float64, np.int64, 2, 8, 5, 1e-8, "Jac", 0.8, 2, 2, 'V')
    mg.MGsetup(A, M, p, 1)
    tile = 256
    for mat, n1, old in ((p.As[0], 41, 5), (p.As[1], 21, 9)):
        monkeypatch.delenv("MGB200_TMA_GAP", raising=False)
        plan = device.host_tma_plan(mat, tile=tile, elem_bytes=elem_bytes)
        assert plan is not None and len(plan["windows"]) == 3
        _check_plan(plan, tile, al)
        # the middle window spans the three lines of the own plane
        lo, ln, _ = plan["windows"][1]
        assert lo <= -(n1 + 1) and lo + ln - tile >= n1 + 1
        monkeypatch.setenv("MGB200_TMA_GAP", "8")
        plan_old = device.host_tma_plan(mat, tile=tile, elem_bytes=elem_bytes)
        assert len(plan_old["windows"]) == old
        _check_plan(plan_old, tile, al)
        assert plan["total"] < plan_old["total"]        # fewer elements copied per tile


def test_tma_plan_absent_for_transfer_operators():
    """P and R take their offsets from the first stored column, not from the row: no TMA plan."""
    from multigrid_jl_b200 import device
    A, AT, M, p, b = _cpu_problem([16, 16], 3)
    assert device.host_tma_plan(p.Ps[0]) is None
    assert device.host_tma_plan(p.As[0], tile=64) is not None


@pytest.mark.parametrize("n,S,S2", [([24, 20, 12], 25, 525), ([32, 32], 33, 0), ([18, 40, 10], 19, 779)])
def test_box_structure_detected(n, S, S2):
    """The line length and the plane length come out of the dictionary offsets alone (7- / 5-point fine level, 27- /
    9-point Galerkin level); a pattern's mask has one bit per stored entry."""
    from multigrid_jl_b200 import device
    A, AT, M, p, b = _cpu_problem(n, 2)
    for l, (s1, s2) in enumerate([(S, S2), (n[0] // 2 + 1, (n[0] // 2 + 1) * (n[1] // 2 + 1) if len(n) == 3 else 0)]):
        mat = sp.csc_matrix(p.As[l])
        box = device.host_detect_box(mat)
        pat = device.host_build_patterns(mat)
        assert box is not None and (box["S"], box["S2"]) == (s1, s2), (l, box)
        lens = np.diff(pat["pat_off"])
        assert [bin(m).count("1") for m in box["masks"]] == list(lens)
        # the centre bit (dz = dy = dx = 0) is set in every pattern of these operators
        assert all(m & (1 << 13) for m in box["masks"])
        # interior pattern: all 27 (9 in 2-D) neighbours on the Galerkin level, 7 (5) on the fine level
        full = max(bin(m).count("1") for m in box["masks"])
        assert full == ((27 if len(n) == 3 else 9) if l == 1 else (7 if len(n) == 3 else 5))


def test_box_structure_rejected():
    from multigrid_jl_b200 import device
    A, AT, M, p, b = _cpu_problem([16, 16], 3)
    assert device.host_detect_box(p.Ps[0]) is None            # not row-relative
    n = 400
    # a pentadiagonal matrix IS a box stencil for lines of 3 (+-2 = +-3 -+ 1): the structure is about offsets only
    T = sp.diags([np.ones(n - 2), np.ones(n - 1), 4 * np.ones(n), np.ones(n - 1), np.ones(n - 2)], [-2, -1, 0, 1, 2], format="csc")
    box = device.host_detect_box(T)
    assert box is not None and (box["S"], box["S2"]) == (3, 0)
    # balanced base-S digits represent many offset sets; {0, +-1, +-3, +-6, +-9, +-30} fits no (S, S2)
    offs = [-30, -9, -6, -3, -1, 0, 1, 3, 6, 9, 30]
    T = sp.diags([np.ones(n - abs(o)) * (4.0 if o == 0 else 1.0) for o in offs], offs, format="csc")
    assert device.host_detect_box(T) is None
    T3 = sp.diags([np.ones(n - 1), 4 * np.ones(n), np.ones(n - 1)], [-1, 0, 1], format="csc")
    box = device.host_detect_box(T3)
    assert box is not None and (box["S"], box["S2"]) == (0, 0)


@pytest.mark.parametrize("n", [[24, 20, 12], [18, 40, 10], [8, 8, 8], [7, 9, 5]])
@pytest.mark.parametrize("RZ,NB", [(1, 256), (1, 64), (2, 128), (2, 40), (4, 64), (4, 512)])
def test_box_kernel_code_on_the_cpu(n, RZ, NB):
    """csrc/box.cuh: the box-stencil kernel's tile plan, copy list and per-thread function are __host__ __device__.
    Replayed on the CPU for every tile of a launch (stages = host buffers filled where the bulk copies fill shared
    memory, zero elsewhere) they must reproduce the one-row-per-thread dictionary walk bit for bit in all modes, on the
    7-point fine level and the 27-point Galerkin level, for tiles that do and do not divide the planes."""
    from multigrid_jl_b200 import device
    A, AT, M, p, b0 = _cpu_problem(n, 2)
    rng = np.random.default_rng(13)
    for l in range(2):
        mat = sp.csc_matrix(p.As[l])
        N = mat.shape[0]
        x, b = rng.standard_normal(N), rng.standard_normal(N)
        d = np.ascontiguousarray(p.relaxPrecs[l]) if l < len(p.relaxPrecs) else 0.8 / mat.diagonal()
        for mode, fold in ((0, False), (2, False), (3, False), (3, True)):
            ref = device.host_pattern_apply(mat, mode, x, b, d, fold)
            got = device.host_box_apply(mat, mode, RZ, NB, x, b, d, fold)
            if ref is None:          # tiny coarse levels are not worth a dictionary
                assert got is None and N < 8 * 27
                continue
            assert got is not None, (l, mode)
            assert got[1]["shape"] == (7 if l == 0 else 27)
            assert np.array_equal(ref[0].view(np.int64), got[0].view(np.int64)), (l, mode, fold)
        if ref is not None:
            # MODE 4: the first two sweeps from x = 0 in one pass == diag_scale followed by one sweep, bit for bit
            dfold = device.host_pattern_apply(mat, 3, np.zeros(N), b, d, True)     # x1 = 0 + d .* (b - A 0): folded d
            x1 = 0.0 + (dfold[0] / np.where(b != 0, b, 1.0)) * b                      # = dpat[pid] .* b
            two = device.host_pattern_apply(mat, 3, x1, b, d, True)
            got4 = device.host_box_apply(mat, 4, RZ, NB, b, b, d, True)
            assert got4 is not None and np.array_equal(two[0].view(np.int64), got4[0].view(np.int64)), l
        if min(n) >= 8 and l == 0 and got is not None:
            assert got[1]["fast_rows"] > 0.2 * N       # interior rows take the constant-coefficient path


def test_box_kernel_rejects_what_it_cannot_do():
    from multigrid_jl_b200 import device
    A, AT, M, p, b0 = _cpu_problem([32, 32], 2)          # 2-D: no planes
    x = np.ones(A.shape[0])
    assert device.host_box_apply(sp.csc_matrix(p.As[0]), 0, 1, 64, x) is None
    assert device.host_box_apply(sp.csc_matrix(p.Ps[0]), 0, 1, 64, np.ones(p.Ps[0].shape[1])) is None


@pytest.mark.parametrize("n", [[24, 20, 12], [32, 32], [18, 40, 10], [14, 10, 8]])
@pytest.mark.parametrize("R", [1, 2, 4])
def test_grid_hinted_transfer_kernel_code_on_the_cpu(n, R):
    """csrc/grid_xfer.cuh: with the meshes as a hint the restriction / prolongation kernels read no matrix stream (pattern
    and columns follow from the grid coordinates, verified at upload).  Their per-thread functions are __host__
    __device__: on the CPU they reproduce the dictionary walk bit for bit and agree with scipy; a wrong hint is
    rejected."""
    from multigrid_jl_b200 import device
    A, AT, M, p, b0 = _cpu_problem(n, 2)
    nf, nc = np.asarray(p.Meshes[0].n) + 1, np.asarray(p.Meshes[1].n) + 1
    rng = np.random.default_rng(5)
    Nf, Nc = int(np.prod(nf)), int(np.prod(nc))
    xf, xc, y0 = rng.standard_normal(Nf), rng.standard_normal(Nc), rng.standard_normal(Nf)
    P, Rm = sp.csc_matrix(p.Ps[0]), sp.csc_matrix(p.Rs[0])
    # restriction r_c = R r_f (Rs[l] stores R^T)
    ref = device.host_grid_transfer(Rm, 2, nf, nc, 0, xf, np.zeros(Nc))
    got = device.host_grid_transfer(Rm, 2, nf, nc, R, xf, np.full(Nc, np.nan))
    assert ref is not None and got is not None
    assert np.array_equal(ref.view(np.int64), got.view(np.int64))
    np.testing.assert_allclose(ref, sp.csr_matrix(Rm.T) @ xf, rtol=1e-13, atol=1e-14)
    # prolongation x_f += P x_c (Ps[l] stores P^T)
    ref = device.host_grid_transfer(P, 1, nf, nc, 0, xc, y0)
    got = device.host_grid_transfer(P, 1, nf, nc, R, xc, y0)
    assert ref is not None and got is not None
    assert np.array_equal(ref.view(np.int64), got.view(np.int64))
    np.testing.assert_allclose(ref, y0 + sp.csr_matrix(P.T) @ xc, rtol=1e-13, atol=1e-14)
    # wrong hints: swapped dimensions (when they differ), another matrix
    if len(set(nf.tolist())) > 1:
        assert device.host_grid_transfer(Rm, 2, nf[::-1].copy(), nc[::-1].copy(), R, xf, np.zeros(Nc)) is None
    assert device.host_grid_transfer(P, 2, nf, nc, R, xf, np.zeros(Nc)) is None
    Pbad = P.copy()
    Pbad.data = Pbad.data.copy()
    Pbad.data[7] *= 1.5          # another value is fine (values come from the dictionary) ...
    got = device.host_grid_transfer(Pbad, 1, nf, nc, R, xc, y0)
    refb = device.host_grid_transfer(Pbad, 1, nf, nc, 0, xc, y0)
    assert (got is None and refb is None) or np.array_equal(got.view(np.int64), refb.view(np.int64))


def _cpu_problem(n, levels):
    import multigrid_jl_b200 as mg
    dom = [0.0, 1.0] * len(n)
    M = mg.getRegularMesh(dom, n)
    A = mg.poisson_shifted(M, 1e-4)
    p = mg.getMGparam(np.float64, np.int64, levels, 8, 4, 1e-12, "Jac", 0.8, 2, 2, 'V')
    mg.MGsetup(A, M, p, 1)
    return A, A, M, p, None


# ---- GPU: dictionary kernels vs CSR-stream kernels, graph replay vs eager launches ------------------

def _solve(kind, n, levels, cycle, env):
    import multigrid_jl_b200 as mg
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        A, AT, M, p, b = make_problem(kind, n, levels, cycle=cycle, maxit=5)
        x = np.zeros_like(b)
        x, _, it = mg.solveMG(p, b, x)
        info = [p.device.pattern_info(l + 1, w) for l in range(levels - 1) for w in range(3)]
        res = p.last_resvec.copy()
        p.device.destroy()
        return x, res, it, info
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.gpu
@pytest.mark.parametrize("kind,n,levels", [("poisson", [64, 64], 4), ("poisson", [24, 24, 24], 3),
                                           ("helmholtz", [48, 48], 3), ("helmholtz", [16, 16, 16], 3),
                                           ("diffusion", [24, 24, 12], 3)])
@pytest.mark.parametrize("cycle", ['V', 'W', 'F'])
def test_pattern_and_graph_paths_bit_identical(kind, n, levels, cycle):
    x0, r0, it0, info0 = _solve(kind, n, levels, cycle, {"MGB200_PATTERNS": "0", "MGB200_GRAPHS": "0", "MGB200_TMA": "0"})
    x1, r1, it1, info1 = _solve(kind, n, levels, cycle, {"MGB200_PATTERNS": "1", "MGB200_GRAPHS": "0", "MGB200_TMA": "0"})
    x2, r2, it2, info2 = _solve(kind, n, levels, cycle, {"MGB200_PATTERNS": "1", "MGB200_GRAPHS": "1", "MGB200_TMA": "0"})
    # TMA-staged persistent variant of the dictionary kernel, forced onto these small levels
    x3, r3, it3, info3 = _solve(kind, n, levels, cycle, {"MGB200_PATTERNS": "1", "MGB200_GRAPHS": "1", "MGB200_TMA": "1",
                                                         "MGB200_TMA_MIN_ROWS": "0"})
    assert it3 == it0 and np.array_equal(r0, r3) and np.array_equal(x0, x3)
    assert not any(i["in_use"] for i in info0)
    if kind != "diffusion":
        assert info1[0]["in_use"] and info1[0]["row_relative"] and info1[0]["d_folded"]   # A_1
        assert info1[1]["in_use"] and not info1[1]["row_relative"]                            # P_1
        assert info1[2]["in_use"]                                                             # R_1
    else:
        assert not info1[0]["in_use"]      # variable coefficients: no value dictionary for A
    assert it0 == it1 == it2
    assert np.array_equal(r0, r1) and np.array_equal(r1, r2)
    assert np.array_equal(x0, x1) and np.array_equal(x1, x2)


@pytest.mark.gpu
@pytest.mark.parametrize("kind,n", [("poisson", [40, 36, 28]), ("helmholtz", [33, 31, 17]), ("poisson", [300, 200])])
def test_tma_kernel_with_vector_d_and_krylov(kind, n):
    """TMA variant with relaxPrecs kept as a vector (d tile copied per stage), odd sizes (clamped windows, slack
    elements) and the SpMV / residual modes of the Krylov drivers: identical to the CSR-stream path."""
    import multigrid_jl_b200 as mg

    def run(env):
        old = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        try:
            A, AT, M, p, b = make_problem(kind, n, 3, maxit=12, tol=1e-9)
            x = np.zeros_like(b)
            if kind == "poisson":
                x, _, it = mg.solveCG_MG(AT, p, b, x)
            else:
                x, _, it, _ = mg.solveGMRES_MG(AT, p, b, x, True, 5)
            res = np.array(p.last_resvec, copy=True)
            p.device.destroy()
            return x, res, it
        finally:
            for k, v in old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
    xa, ra, ia = run({"MGB200_PATTERNS": "0", "MGB200_TMA": "0", "MGB200_FOLD_D": "1"})
    xb, rb, ib = run({"MGB200_PATTERNS": "1", "MGB200_TMA": "1", "MGB200_TMA_MIN_ROWS": "0", "MGB200_FOLD_D": "0"})
    xc, rc, ic = run({"MGB200_PATTERNS": "1", "MGB200_TMA": "1", "MGB200_TMA_MIN_ROWS": "0", "MGB200_FOLD_D": "1"})
    assert ia == ib == ic
    assert np.array_equal(ra, rb) and np.array_equal(ra, rc)
    assert np.array_equal(xa, xb) and np.array_equal(xa, xc)


@pytest.mark.gpu
@pytest.mark.parametrize("kind,n,cycle", [("poisson", [40, 36, 28], 'V'), ("helmholtz", [33, 31, 17], 'W'),
                                          ("poisson", [300, 200], 'F')])
@pytest.mark.parametrize("tma", ["0", "1"])
def test_split_launches_bit_identical(kind, n, cycle, tma):
    """The multi-GPU overlap path launches every dictionary pass as interior rows + the rows at both ends
    (pattern_apply_split, launch.cuh).  MGB200_SPLIT_TEST forces that launch sequence on one GPU: the
    results must not change by a bit, for the one-pass kernel and for the TMA-staged tiles."""
    base = {"MGB200_PATTERNS": "1", "MGB200_GRAPHS": "1", "MGB200_TMA": tma, "MGB200_TMA_MIN_ROWS": "0"}
    x0, r0, it0, _ = _solve(kind, n, 3, cycle, dict(base, MGB200_SPLIT_TEST="0"))
    for rows in ("7", "1500"):
        x1, r1, it1, _ = _solve(kind, n, 3, cycle, dict(base, MGB200_SPLIT_TEST=rows))
        assert it0 == it1 and np.array_equal(r0, r1) and np.array_equal(x0, x1)


@pytest.mark.gpu
@pytest.mark.parametrize("kind,n,cycle", [("poisson", [40, 36, 28], 'V'), ("helmholtz", [33, 31, 17], 'W'),
                                          ("poisson", [64, 64, 64], 'F')])
@pytest.mark.parametrize("variant", ["0", "1", "3", "9", "11"])
def test_box_kernel_bit_identical(kind, n, cycle, variant):
    """csrc/box.cuh: the box-stencil kernel (dense coefficient tables, unrolled 7- / 27-point chains, RZ rows per
    thread one plane apart) on every box-structured 3-D level instead of the dictionary walk: results must not change
    by a bit (Float64 and ComplexF64; planes the tiles do and do not divide; d folded and as a vector)."""
    base = {"MGB200_PATTERNS": "1", "MGB200_GRAPHS": "1", "MGB200_TMA": "1", "MGB200_TMA_MIN_ROWS": "0"}
    x0, r0, it0, _ = _solve(kind, n, 3, cycle, dict(base, MGB200_BOX="0"))
    for fold in ("1", "0"):
        x1, r1, it1, _ = _solve(kind, n, 3, cycle, dict(base, MGB200_BOX="1", MGB200_BOX_MIN_ROWS="0",
                                                        MGB200_BOX_VARIANT=variant, MGB200_FOLD_D=fold))
        assert it0 == it1 and np.array_equal(r0, r1) and np.array_equal(x0, x1), fold


@pytest.mark.gpu
@pytest.mark.parametrize("kind,n,cycle", [("poisson", [40, 36, 28], 'V'), ("helmholtz", [32, 24, 16], 'W'),
                                          ("poisson", [300, 200], 'F'), ("poisson", [64, 64, 64], 'V')])
@pytest.mark.parametrize("R", ["1"])
def test_grid_hinted_transfers_bit_identical(kind, n, cycle, R):
    """Option "grid_transfers" (MGB200_GRID_TRANSFERS) + the meshes as a hint at upload: restriction and prolongation
    without a matrix stream.  Results must not change by a bit."""
    base = {"MGB200_PATTERNS": "1", "MGB200_GRAPHS": "1"}
    x0, r0, it0, _ = _solve(kind, n, 3, cycle, dict(base, MGB200_GRID_TRANSFERS="0"))
    x1, r1, it1, _ = _solve(kind, n, 3, cycle, dict(base, MGB200_GRID_TRANSFERS=R))
    assert it0 == it1 and np.array_equal(r0, r1) and np.array_equal(x0, x1)


@pytest.mark.gpu
@pytest.mark.parametrize("n,nrhs", [([24, 20, 18], 32), ([17, 9, 33], 8), ([40, 12, 12], 13), ([16, 16, 16], 40)])
@pytest.mark.parametrize("variant", ["quad_prolongation", "long_lines_30", "long_lines_31", "long_lines_32", "long_lines_33"])
def test_round2_kernel_options_bit_identical(n, nrhs, variant):
    """Marching block kernel (csrc/box.cuh, box_mrhs_march_kernel: tiles that do not divide the grid, nrhs below / at /
    above one warp) against the direct block kernel, and the box variants for long lines / the line-form prolongation
    against the defaults, on one right-hand side: nothing may change by a bit."""
    import multigrid_jl_b200 as mg

    def run(env, m):
        old = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        try:
            A, AT, M, p, b = make_problem("poisson", n, 3, nrhs=m, maxit=3)
            x, _, it = mg.solveMG(p, b, np.zeros_like(b))
            res = p.last_resvec.copy()
            p.device.destroy()
            return x, res
        finally:
            for k, v in old.items():
                os.environ.pop(k, None) if v is None else os.environ.__setitem__(k, v)
    base = {"MGB200_BOX_MIN_ROWS": "0"}
    if variant == "quad_prolongation":
        x0, r0 = run(dict(base, MGB200_MRHS_MARCH="0"), nrhs)
        x1, r1 = run(dict(base, MGB200_MRHS_MARCH="1"), nrhs)
        assert np.array_equal(r0, r1) and np.array_equal(x0, x1)
        for gt in ("0", "2"):      # block transfers: CSR stream (0), grid-hinted restriction (default), + prolongation (2)
            x1, r1 = run(dict(base, MGB200_GRID_TRANSFERS=gt), nrhs)
            assert np.array_equal(r0, r1) and np.array_equal(x0, x1), gt
        x0, r0 = run(dict(base, MGB200_GXP_QUAD="0"), 1)
        for q in ("1", "2", "3"):
            x1, r1 = run(dict(base, MGB200_GXP_QUAD=q), 1)
            assert np.array_equal(r0, r1) and np.array_equal(x0, x1), q
    else:
        x0, r0 = run(dict(base, MGB200_BOX="0"), 1)
        x1, r1 = run(dict(base, MGB200_BOX_VARIANT=variant.split("_")[-1]), 1)
        assert np.array_equal(r0, r1) and np.array_equal(x0, x1)


@pytest.mark.parametrize("seed", range(12))
def test_box_kernel_code_fuzz(seed):
    """Random box stencils - arbitrary subsets of the 27 offsets (upwind-like one-sided ones, stencils without a centre
    entry, different ones near the boundaries), random grid sizes with tiny dimensions: the box-stencil kernel code
    (csrc/box.cuh, replayed on the CPU) must equal the dictionary walk bit for bit whenever the matrix qualifies."""
    from multigrid_jl_b200 import device
    rng = np.random.default_rng(1000 + seed)
    dim = 3 if seed % 3 else 2
    n = [int(rng.integers(4, 14)) for _ in range(dim)] + [1] * (3 - dim)
    n[0] = int(rng.integers(5, 40))
    N = n[0] * n[1] * n[2]
    S, S2 = n[0], n[0] * n[1]
    offs = [(dz, dy, dx) for dz in ((-1, 0, 1) if dim == 3 else (0,)) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]
    keep = [o for o in offs if rng.random() < 0.6]
    if not any(o[2] != 0 for o in keep):
        keep.append((0, 0, 1))
    if not any(o[1] != 0 for o in keep):
        keep.append((0, 1, 0))
    vals = {o: float(rng.standard_normal()) for o in keep}
    rows, cols, data = [], [], []
    idx = np.arange(N)
    i, j, k = idx % n[0], (idx // n[0]) % n[1], idx // S2
    for (dz, dy, dx) in keep:
        ok = (i + dx >= 0) & (i + dx < n[0]) & (j + dy >= 0) & (j + dy < n[1]) & (k + dz >= 0) & (k + dz < n[2])
        # a second value on the first plane / line makes more patterns than the pure boundary classes
        v = np.where((j == 0) | (k == 0), vals[(dz, dy, dx)] * 1.5, vals[(dz, dy, dx)])
        rows.append(idx[ok]); cols.append(idx[ok] + dx + S * dy + S2 * dz); data.append(v[ok])
    A = sp.csr_matrix((np.concatenate(data), (np.concatenate(rows), np.concatenate(cols))), shape=(N, N))
    A.sort_indices()
    mat = sp.csc_matrix(A.T)          # the hierarchy stores the transpose (CSC arrays = CSR arrays of the operator)
    x, b, d = rng.standard_normal(N), rng.standard_normal(N), rng.standard_normal(N)
    ref = device.host_pattern_apply(mat, 3, x, b, d, False)
    if ref is None:
        # no dictionary (too few rows per pattern) or an offset set that fits no (S, S2): nothing to compare
        return
    np.testing.assert_allclose(ref[0], x + d * (b - A @ x), rtol=1e-12, atol=1e-12)
    for RZ, NB in ((1, 64), (2, 40), (4, 128)):
        for mode in (0, 2, 3):
            r0 = device.host_pattern_apply(mat, mode, x, b, d, False)
            got = device.host_box_apply(mat, mode, RZ, NB, x, b, d, False)
            if got is None:          # 2-D grids, or a box the kernel is not built for
                assert dim == 2 or device.host_detect_box(mat) is None or device.host_detect_box(mat)["S2"] < 3 * device.host_detect_box(mat)["S"]
                continue
            assert np.array_equal(r0[0].view(np.int64), got[0].view(np.int64)), (n, keep, RZ, NB, mode)


def test_box_structure_of_a_z_slab_is_read_the_simple_way():
    """A z-slab of 16 planes of 17 x 17 nodes (rank 0 of a row-partitioned level) tiles as lines of 16 as well:
    offsets +-17 = (dy,dx) = (1,1), +-289 = (dz,dy,dx) = (1,1,1) with S2 = 272, and 16 * 289 = 17 * 272 rows.  The
    detection must still read the 7-point stencil as a star (S = 17, S2 = 289): the other reading runs the 27-point
    kernel on it (found on the B200: 150 us against 97 us per level-1 sweep on rank 0, profiles/r02o_*)."""
    from multigrid_jl_b200 import device
    import multigrid_jl_b200 as mg
    n, dom = [16, 16, 32], [0, 1, 0, 1, 0, 2]
    A = mg.poisson_shifted(mg.getRegularMesh(dom, n), 1e-4).tocsr()
    plane = 17 * 17
    rows = A[:16 * plane, :].tocoo()
    loc = sp.csr_matrix((rows.data, (rows.row, rows.col)), shape=(16 * plane, 17 * plane))     # owned rows | ghost plane above
    box = device.host_detect_box(sp.csc_matrix(loc.T))
    assert box is not None and (box["S"], box["S2"]) == (17, 289)
    star = (1 << 4) | (1 << 10) | (1 << 12) | (1 << 13) | (1 << 14) | (1 << 16) | (1 << 22)
    assert all((int(m) & ~star) == 0 for m in box["masks"])
