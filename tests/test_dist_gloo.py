"""N > 1 host logic on CPU: 2 (and 3) ranks over gloo - slab setup (windowed Galerkin), ghost
planning through the library's host-only planner, halo exchange / coarse all-gather / norm
all-reduce pattern - must reproduce the global oracle's per-cycle residual norms to 1e-10."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 3])
def test_distributed_cycle_matches_global_oracle(world):
    port = _free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_dist_worker.py")], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r} failed:\n{o[-3000:]}"
    assert "DIST_OK" in outs[0]


def test_slab_planes_follow_the_reference_box_partition():
    """getOriginalBoundingBoxCells (DDIndices.jl:41-47): cellSize = div(nc, NumCells), last box absorbs
    the remainder; node planes: slab g owns g*c .. (g+1)*c-1, the last slab also the final plane."""
    import multigrid_jl_b200 as mg
    assert mg.slab_planes(256, 8) == [(32 * g, 32 * (g + 1)) for g in range(7)] + [(224, 257)]
    assert mg.slab_planes(10, 3) == [(0, 3), (3, 6), (6, 11)]
    with pytest.raises(ValueError):
        mg.slab_planes(2, 4)


@pytest.mark.parametrize("n,world,nd_expected", [([8, 8, 64], 4, 3), ([8, 16, 48], 3, 3), ([4, 4, 96], 8, 3),
                                                 ([4, 8, 40], 3, 2), ([8, 8, 32], 2, 2)])
def test_windowed_galerkin_equals_global_rows(n, world, nd_expected):
    """Every rank's owned rows (A, P, R, d) of every distributed level equal the global hierarchy bit for bit
    (also when the slab boundaries of different levels do not coincide and for slabs a few planes thin)."""
    import scipy.sparse as sp
    import multigrid_jl_b200 as mg
    dom = [0, 1, 0, 1, 0, n[2] / 16.0]      # dyadic mesh widths: window and global operators agree to the bit
    M = mg.getRegularMesh(dom, n)
    pg = mg.getMGparam(np.float64, np.int64, 5, 8, 5, 1e-8, 'SPAI', 1.0, 2, 2, 'V')
    mg.MGsetup(mg.poisson_shifted(M, 1e-4), M, pg, 1)
    op = mg.poisson_window_operator(dom, n, 1e-4)
    store = {}

    def run(rank, pieces):
        p = mg.getMGparam(np.float64, np.int64, 5, 8, 5, 1e-8, 'SPAI', 1.0, 2, 2, 'V')

        def gather(o):
            if pieces is None:
                store[rank] = o
                raise StopIteration
            return pieces
        try:
            return mg.setup_slab_hierarchy(op, dom, n, p, rank, world, replicate_below=100, gather=gather)
        except StopIteration:
            return None
    for r in range(world):
        run(r, None)
    dhs = [run(r, [store[q] for q in range(world)]) for r in range(world)]
    nd = dhs[0].nd
    assert nd == nd_expected
    for l in range(nd):
        for name, ref in (("AT", pg.As[l]), ("PT", pg.Ps[l]), ("RT", pg.Rs[l])):
            glob = sp.hstack([getattr(dh.dist_levels[l], name) for dh in dhs]).tocsc()
            assert (glob != ref).nnz == 0, (l, name)
        assert np.array_equal(np.concatenate([dh.dist_levels[l].d for dh in dhs]), pg.relaxPrecs[l])
    for j, a in enumerate(dhs[0].replicated.As):
        assert (a != pg.As[nd + j]).nnz == 0


@pytest.mark.parametrize("n,world", [([8, 8, 64], 4), ([8, 16, 48], 3), ([4, 4, 96], 8), ([8, 8, 32], 2), ([8, 8, 16], 1)])
def test_rediscretised_slabs_equal_global_rows(n, world):
    """cfg5's hierarchy (ComplexF64 shifted Laplacian, rediscretised on every level: the multilevelOperatorConstructor
    path, MGsetup.jl:28,105-106): every rank's owned rows of A, P, R and d equal the global MGsetup bit for bit, and
    the replicated coarse part equals the global coarse levels."""
    import scipy.sparse as sp
    import multigrid_jl_b200 as mg
    dom = [0, 1, 0, n[1] / n[0], 0, n[2] / n[0]]
    kappa2 = (2 * np.pi / (10 * (1.0 / n[0])) * 0.35) ** 2
    op = mg.poisson_window_operator(dom, n, kappa2=kappa2, gamma=0.5)
    M = mg.getRegularMesh(dom, n)
    ctor = mg.getMultilevelOperatorConstructor(kappa2, lambda mesh, k2: op(np.asarray(mesh.n), 0, int(mesh.n[2]) + 1),
                                               lambda mf, mc, pf, level: pf)
    pg = mg.getMGparam(np.complex128, np.int64, 5, 8, 5, 1e-8, 'Jac', 0.8, 2, 2, 'V')
    mg.MGsetup(ctor, M, pg, 1)
    dhs = []
    for r in range(world):
        p = mg.getMGparam(np.complex128, np.int64, 5, 8, 5, 1e-8, 'Jac', 0.8, 2, 2, 'V')
        dhs.append(mg.setup_slab_hierarchy(op, dom, n, p, r, world, replicate_below=100, gather=lambda o: [o],
                                           rediscretise=True))
    nd = dhs[0].nd
    assert nd >= 1 or world == 1
    for l in range(nd):
        for name, ref in (("AT", pg.As[l]), ("PT", pg.Ps[l]), ("RT", pg.Rs[l])):
            glob = sp.hstack([getattr(dh.dist_levels[l], name) for dh in dhs]).tocsc()
            assert glob.shape == ref.shape and (glob != ref).nnz == 0, (l, name)
        d = np.concatenate([dh.dist_levels[l].d for dh in dhs])
        assert np.array_equal(d, pg.relaxPrecs[l])
    rep = dhs[0].replicated
    assert len(rep.As) == len(pg.As) - nd
    for j in range(len(rep.As)):
        assert (rep.As[j] != pg.As[nd + j]).nnz == 0
