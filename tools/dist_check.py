#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/dist_check.py

Every rank uploads its z-slab of the hierarchy; the row-partitioned GPU path (NCCL halo
exchange, coarse all-gather, scalar all-reduces) must reproduce the GLOBAL CPU oracle:
per-cycle residual norms of solveMG within 1e-10, same PCG / FGMRES iteration counts.
Rank 0 prints one JSON line per case and exits non-zero on a mismatch."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import multigrid_jl_b200 as mg
    from oracle import cycle as oc

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("gloo")

    def gather(o):
        out = [None] * world
        dist.all_gather_object(out, o)
        return out
    ok = True
    cases = [("poisson", [32, 32, 32 * world], 5, 'V', np.float64), ("poisson", [32, 32, 32 * world], 5, 'W', np.float64),
             ("helmholtz", [32, 32, 32 * world], 4, 'V', np.complex128), ("poisson", [16, 16, 16 * world], 4, 'K', np.float64),
             # thin slabs of wide planes (the shape of the 512^3 weak-scaling point: few planes per GPU)
             ("poisson", [32, 64, 8 * world], 4, 'V', np.float64)]
    only = os.environ.get("DIST_CHECK_CASES")
    if only:
        cases = [cases[int(i)] for i in only.split(",")]
    for kind, n, levels, cyc, VAL in cases:
        dom = [0, 1, 0, n[1] / n[0], 0, n[2] / n[0]]
        p = mg.getMGparam(VAL, np.int64, levels, 8, 5, 1e-12, "Jac", 0.8, 2, 2, cyc)
        p.nrhs = 1
        h = 1.0 / n[0]
        kappa2 = (2 * np.pi / (10 * h) * 0.35) ** 2
        if kind == "poisson":
            op = mg.poisson_window_operator(dom, n, 1e-4)
        else:
            op = mg.poisson_window_operator(dom, n, kappa2=kappa2, gamma=0.5)
        dh = mg.setup_slab_hierarchy(op, dom, n, p, rank, world, replicate_below=3000, gather=gather)
        ids = [mg.DeviceHierarchy.dist_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        dev = mg.DeviceHierarchy.from_dist(dh, p, local_rank, ids[0])
        # global problem + oracle on rank 0 (every rank needs b)
        Mg = mg.getRegularMesh(dom, n)
        Ag = mg.poisson_shifted(Mg, 1e-4) if kind == "poisson" else mg.helmholtz_shifted(Mg, kappa2, 0.5)
        rng = np.random.default_rng(0)
        u = rng.random(Ag.shape[0]) + (1j * rng.random(Ag.shape[0]) if VAL == np.complex128 else 0)
        bg = (Ag @ u).astype(VAL)
        bg /= np.linalg.norm(bg)
        lo, hi = int(dh.dist_levels[0].row_offsets[rank]), int(dh.dist_levels[0].row_offsets[rank + 1])
        b = np.ascontiguousarray(bg[lo:hi])
        x, it, res = dev.solveMG(b, np.zeros_like(b), 1e-12, 5)
        rec = {"case": f"{kind} {n} {cyc} levels={levels}", "world": world, "dist_levels": dh.nd}
        if rank == 0:
            pg = mg.getMGparam(VAL, np.int64, levels, 8, 5, 1e-12, "Jac", 0.8, 2, 2, cyc)
            ATg = Ag.conj().T.tocsc() if VAL == np.complex128 else Ag
            mg.MGsetup(ATg, Mg, pg, 1)
            o = oc.OracleMG(pg)
            xr, itr, res_ref = oc.solveMG(o, bg, np.zeros_like(bg))
            err = float(np.max(np.abs(res - res_ref) / res_ref))
            rec.update(solveMG_iter=[it, itr], solveMG_maxrel=err)
            ok &= (it == itr and err < 1e-10)
        xs = gather(x)
        if rank == 0:
            xe = float(np.linalg.norm(np.concatenate(xs) - xr) / np.linalg.norm(xr))
            rec.update(x_relerr=xe)
            ok &= xe < 1e-9
        # Krylov: PCG (real) or FGMRES (complex)
        p.relativeTol, p.maxOuterIter = 1e-8, 30
        if VAL == np.float64 and cyc == 'V':
            xk, itk, flag, resv = dev.solveCG(b, np.zeros_like(b), 1e-8, 30)
            if rank == 0:
                pg.relativeTol, pg.maxOuterIter = 1e-8, 30
                o = oc.OracleMG(pg)
                xr2, itr2, flagr, resr = oc.solveCG_MG(ATg, o, bg, np.zeros_like(bg))
                rec.update(cg_iter=[itk, itr2], cg_maxrel=float(np.max(np.abs(resv - resr) / resr)))
                ok &= (itk == itr2 and flag == flagr and rec["cg_maxrel"] < 1e-7)
        if VAL == np.complex128:
            xk, itk, flag, resv = dev.solveFGMRES(b, np.zeros_like(b), 5, True, 1e-8, 10)
            if rank == 0:
                pg.relativeTol, pg.maxOuterIter = 1e-8, 10
                o = oc.OracleMG(pg)
                xr2, itr2, flagr, resr = oc.solveGMRES_MG(ATg, o, bg, np.zeros_like(bg), True, 5)
                rec.update(fgmres_iter=[itk, itr2], fgmres_nres=[len(resv), len(resr)],
                           fgmres_maxrel=float(np.max(np.abs(resv - resr[:len(resv)]) / resr[:len(resv)])))
                ok &= (itk == itr2 and len(resv) == len(resr) and rec["fgmres_maxrel"] < 1e-6)
        if rank == 0:
            print(json.dumps(rec), flush=True)
        dev.destroy()
        dist.barrier()
    flag = torch.tensor([1 if ok else 0])
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    if rank == 0:
        print("DIST_CHECK", "PASS" if ok else "FAIL", flush=True)
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
