"""Worker of tests/test_dist_gloo.py: world_size ranks over gloo on CPU.

Each rank builds its slab of the hierarchy (multigrid_jl_b200.dist_setup), plans its ghost layout with
the library's host-only planner (mgb200_host_plan_ghosts, the same code the GPU path uses), and runs
the distributed V-cycle in the reference's operation order with scipy on the local [owned|ghost]
data, exchanging halos / all-gathering the coarse right-hand side / all-reducing norms through
gloo - the same communication pattern solver.cuh issues over NCCL.  Rank 0 compares the per-cycle
residual norms with the global CPU oracle."""
import ctypes
import os
import sys

import numpy as np
import scipy.sparse as sp
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import multigrid_jl_b200 as mg  # noqa: E402
from multigrid_jl_b200 import device  # noqa: E402


def plan(cols_list, lo, hi):
    """ghost set of the union of several global-column arrays + remapped copies (host ABI)."""
    L = device.lib()
    allc = np.ascontiguousarray(np.concatenate(cols_list), dtype=np.int64)
    ghosts = np.zeros(max(allc.size, 1), dtype=np.int64)
    ng = ctypes.c_int64(0)
    loc = np.zeros(max(allc.size, 1), dtype=np.int64)
    st = L.mgb200_host_plan_ghosts(ctypes.c_int64(allc.size), allc.ctypes.data_as(ctypes.c_void_p),
                                   ctypes.c_int64(lo), ctypes.c_int64(hi), ghosts.ctypes.data_as(ctypes.c_void_p),
                                   ctypes.byref(ng), loc.ctypes.data_as(ctypes.c_void_p))
    assert st == 0
    out, o = [], 0
    for c in cols_list:
        out.append(loc[o:o + c.size].copy())
        o += c.size
    return ghosts[:ng.value].copy(), out


class Space:
    def __init__(self, lo, hi, ghosts, row_offsets, rank, world):
        self.lo, self.hi, self.ghosts = lo, hi, ghosts
        self.n_owned = hi - lo
        owners = np.searchsorted(row_offsets, ghosts, side="right") - 1
        self.recv = [ghosts[owners == p] for p in range(world)]
        reqs = [None] * world
        dist.all_gather_object(reqs, self.recv)
        self.send = [np.asarray(reqs[q][rank], dtype=np.int64) - lo for q in range(world)]
        self.rank, self.world = rank, world

    def exchange(self, v):
        """v: [owned | ghost]; fills the ghost part."""
        packs = [v[self.send[q]] for q in range(self.world)]
        got = [None] * self.world
        dist.all_gather_object(got, packs)
        o = self.n_owned
        for p in range(self.world):
            piece = got[p][self.rank]
            v[o:o + len(piece)] = piece
            o += len(piece)


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, dom, levels = [8, 8, 32], [0, 1, 0, 1, 0, 2.0], 4
    param = mg.getMGparam(np.float64, np.int64, levels, 8, 4, 1e-12, "Jac", 0.8, 2, 2, 'V')

    def gather(o):
        out = [None] * world
        dist.all_gather_object(out, o)
        return out
    dh = mg.setup_slab_hierarchy(mg.poisson_window_operator(dom, n, 1e-4), dom, n, param, rank, world,
                                 replicate_below=300, gather=gather)
    nd = dh.nd
    assert nd >= 2, "test wants at least two distributed levels"
    rep = dh.replicated
    # ---- local layouts -------------------------------------------------------------------------
    spaces, Aloc, Ploc, Rloc = [], [], [], []
    cols = []
    for l, dl in enumerate(dh.dist_levels):
        c = [dl.AT.indices.astype(np.int64), dl.RT.indices.astype(np.int64)]
        if l > 0:
            c.append(dh.dist_levels[l - 1].PT.indices.astype(np.int64))
        lo, hi = int(dl.row_offsets[rank]), int(dl.row_offsets[rank + 1])
        ghosts, loc = plan(c, lo, hi)
        spaces.append(Space(lo, hi, ghosts, dl.row_offsets, rank, world))
        cols.append(loc)
    for l, dl in enumerate(dh.dist_levels):
        s = spaces[l]
        nloc = s.n_owned + len(s.ghosts)
        Aloc.append(sp.csr_matrix((dl.AT.data, cols[l][0], dl.AT.indptr), shape=(s.n_owned, nloc)))
        nco = int(dl.coarse_row_offsets[rank + 1] - dl.coarse_row_offsets[rank])
        Rloc.append(sp.csr_matrix((dl.RT.data, cols[l][1], dl.RT.indptr), shape=(nco, nloc)))
        if l + 1 < nd:
            pc = cols[l + 1][2]
            ncl = spaces[l + 1].n_owned + len(spaces[l + 1].ghosts)
        else:
            pc = dl.PT.indices.astype(np.int64)
            ncl = dl.nc_global
        Ploc.append(sp.csr_matrix((dl.PT.data, pc, dl.PT.indptr), shape=(s.n_owned, ncl)))
    # replicated part: the global oracle cycle on levels nd..L
    from oracle import cycle as oc
    orep = oc.OracleMG(rep, numCores=1)
    orep.relaxPre = lambda l: param.relaxPre(l + nd)
    orep.relaxPost = lambda l: param.relaxPost(l + nd)
    orep.adjustMemoryForNumRHS(1)

    def gnorm(v_owned):
        t = torch.tensor([float(np.dot(v_owned, v_owned))], dtype=torch.float64)
        dist.all_reduce(t)
        return float(np.sqrt(t.item()))

    def cycle(l, b, x, xzero):
        s, A, d = spaces[l], Aloc[l], dh.dist_levels[l].d
        no = s.n_owned
        pre, post = param.relaxPre(l + 1), param.relaxPost(l + 1)

        def resid():
            s.exchange(x)
            return b - A @ x
        r = b.copy() if xzero else resid()
        for i in range(pre):
            if i > 0:
                r = resid()
            x[:no] += d * r
        r = resid()
        rfull = np.zeros_like(x)
        rfull[:no] = r
        s.exchange(rfull)
        bc_own = Rloc[l] @ rfull
        if l + 1 < nd:
            sc = spaces[l + 1]
            xc = np.zeros(sc.n_owned + len(sc.ghosts))
            cycle(l + 1, bc_own, xc, True)
            sc.exchange(xc)
        else:
            pieces = [None] * world
            dist.all_gather_object(pieces, bc_own)
            bc = np.concatenate(pieces)
            xc = np.zeros_like(bc)
            xc = oc.recursiveCycle(orep, bc, xc, 1)
        x[:no] += Ploc[l] @ xc
        for i in range(post):
            r = resid()
            x[:no] += d * r
        return x

    # ---- distributed solveMG ------------------------------------------------------------------------
    rng = np.random.default_rng(0)
    Mg = mg.getRegularMesh(dom, n)
    Ag = mg.poisson_shifted(Mg, 1e-4)
    bg = Ag @ rng.random(Ag.shape[0])
    bg /= np.linalg.norm(bg)
    s0 = spaces[0]
    b = bg[s0.lo:s0.hi].copy()
    x = np.zeros(s0.n_owned + len(s0.ghosts))
    res = [gnorm(b)]
    for it in range(param.maxOuterIter):
        x = cycle(0, b, x, it == 0)
        s0.exchange(x)
        res.append(gnorm(b - Aloc[0] @ x))
    if rank == 0:
        pg = mg.getMGparam(np.float64, np.int64, levels, 8, 4, 1e-12, "Jac", 0.8, 2, 2, 'V')
        mg.MGsetup(Ag, Mg, pg, 1)
        _, itr, res_ref = oc.solveMG(oc.OracleMG(pg, numCores=1), bg, np.zeros_like(bg))
        err = np.max(np.abs(np.array(res) - res_ref) / res_ref)
        print(f"DIST_OK nd={nd} ghosts={[len(s.ghosts) for s in spaces]} maxrel={err:.3e}", flush=True)
        assert err < 1e-10, (res, res_ref)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
