// Device-resident multigrid cycle and Krylov drivers.
//
//   Hierarchy::cycle            <- recursiveCycle           src/Multigrid/MGcycle.jl:1-118
//   Hierarchy::relax            <- relax                    src/Multigrid/MGcycle.jl:122-136
//   Hierarchy::solve_coarsest   <- solveCoarsest            src/Multigrid/MGcycle.jl:138-181
//   Hierarchy::fgmres_relaxation<- FGMRES_relaxation        src/Multigrid/FGMRES.jl:48-126
//   Hierarchy::solveMG          <- solveMG                  src/Multigrid/SolveFuncs.jl:3-39
//   Hierarchy::solveCG          <- solveCG_MG -> KrylovMethods.cg      SolveFuncs.jl:103-116
//   Hierarchy::solveFGMRES      <- solveGMRES_MG -> KrylovMethods.fgmres SolveFuncs.jl:120-132
//
// The control flow runs on the host and only enqueues kernels; the V/F/W cycles contain no
// host synchronisation at all (the reference's data-dependent `if norm(x)>0` is replaced by a
// host-tracked "x is known to be zero" flag, which is result-identical: r - A*0 == r).
#pragma once
#include <functional>
#include <map>
#include <tuple>

#include "dist.cuh"
#include "launch.cuh"
#include "p2p.cuh"
#include "smalldense.h"

namespace mgb200 {

template <typename TV>
struct FgmresMem {  // FGMRESmem (FGMRES.jl:3-8): Z and V(=AZ) are (n*m) x inner
    TV* Z = nullptr;
    TV* AZ = nullptr;
    TV* Az = nullptr;   // result buffer of Afun
    TV* vp0 = nullptr;  // v_prec (x of the recursive call / D.*v) ...
    TV* vp1 = nullptr;  // ... and its ping-pong partner
    int inner = 0;
    void release() {
        dev_free(Z); dev_free(AZ); dev_free(Az); dev_free(vp0); dev_free(vp1);
        inner = 0;
    }
};

template <typename TV>
struct Level {
    long long n = 0;
    Csr<TV> A;
    typedef typename VT<TV>::real_t RT;   // P and R are real (SA-AMG.jl:9-10, MGsetup.jl:80-81)
    Csr<RT> P, R;
    TV* d = nullptr;
    TV* dpat = nullptr;            // d as a function of A's pattern id, when it is one (pattern.cuh)
    int gn[3] = {0, 0, 0}, gN[3] = {0, 0, 0};   // grid hint (mgb200_set_level_grid): fine / coarse nodes per dimension
    bool ghint = false;
    TV *b = nullptr, *r = nullptr, *x0 = nullptr, *x1 = nullptr;  // CYCLEmem + ping-pong partner of x
    FgmresMem<TV> memRelax, memK;
    // ---- multi-GPU row partition (dist.cuh) ----
    long long nalloc = 0;          // rows allocated per vector: owned + ghost (== n when not distributed)
    DistSpace sp;                  // vector space of this level
    HostRows<TV> hA;               // staged owned rows with global columns (until finalisation)
    HostRows<RT> hP, hR;
    std::vector<TV> hd;
    std::vector<long long> coarse_row_offsets;  // row partition of level l+1 (R output / P input)
    long long nc_global = 0;
    void release_work() {
        dev_free(b); dev_free(r); dev_free(x0); dev_free(x1);
        memRelax.release();
        memK.release();
    }
    void release() {
        release_work();
        A.release(); P.release(); R.release();
        dev_free(d);
        dev_free(dpat);
        sp.release();
        n = 0;
    }
};

template <typename TV>
struct Coarsest {
    int n = 0;
    int kind = 0;     // 0: dense LU (default branch of defineCoarsestAinv), 1: "GMRES" (MGsetup.jl:333-334)
    typedef typename Wide<TV>::type TW;   // the factors of a single-precision hierarchy are double precision (common.cuh)
    TW* linv = nullptr;
    TW* uinv = nullptr;
    int* perm = nullptr;
    TW* y = nullptr;  // n*m scratch
    // coarseSolveType "GMRES": param.LU = conj(relaxParam ./ diag(AT)) and the FGMRES(10) workspace
    TV *d = nullptr, *r = nullptr, *w = nullptr, *t = nullptr, *V = nullptr, *dv = nullptr;
    void release() {
        dev_free(linv); dev_free(uinv); dev_free(perm); dev_free(y);
        dev_free(d); dev_free(r); dev_free(w); dev_free(t); dev_free(V); dev_free(dv);
        n = 0;
        kind = 0;
    }
};

struct HierarchyBase {
    virtual ~HierarchyBase() {}
    int val_type = 0;
    Context ctx;
};

template <typename TV>
struct Hierarchy : HierarchyBase {
    typedef typename VT<TV>::real_t RT;
    int levels = 0;
    int m = 1;  // nrhs
    char cycle_type = 'V';
    int relax_kind = 0;
    std::vector<int> pre, post;
    std::vector<Level<TV>> L;  // L[0] finest ... L[levels-1] coarsest (only n, b, r, x used there)
    Coarsest<TV> coarse;
    Csr<TV> Akry;              // optional Krylov matrix
    bool work_ready = false;
    // Krylov workspace (fine level)
    TV *kr = nullptr, *kp = nullptr, *kAp = nullptr, *kw = nullptr, *kV = nullptr, *kZ = nullptr;
    int kV_cols = 0;
    TV* hstage = nullptr;  // device staging for host<->device layout changes
    long long hstage_n = 0;
    // the caller's x (level-1 iterate of solveMG / the Krylov iterate) and its ping-pong partner;
    // L[0].x0/x1 are memCycle[1].x, the z of the preconditioner closure (SolveFuncs.jl:50,59)
    TV *ux0 = nullptr, *ux1 = nullptr, *ucur = nullptr;
    Comm comm;                 // NCCL communicator (world == 1: inactive)
    P2P p2p;                   // halo exchange / coarse gather over NVLink peer memory (p2p.cuh)
    // fused put (ll.cuh): put_done[l] = the level-l vector whose slab-end rows its producing kernel has already
    // stored to the neighbours for the NEXT exchange of channel l; that exchange then only polls and unpacks
    std::vector<const TV*> put_done;
    int use_fused_put = 1;
    bool dist_finalized = true;
    // mixed precision (mgb200_create_mixed): a handle without levels of its own whose preconditioner is one cycle of
    // a single-precision hierarchy (getMultigridPreconditioner with VAL != eltype(B), SolveFuncs.jl:52-60)
    bool krylov_only = false;
    std::function<TV*(const TV*)> ext_prec;
    cudaEvent_t mix_ev[2] = {nullptr, nullptr};
    // V/F/W cycles have no host synchronisation: each (buffers, x-is-zero, type) variant is captured
    // once into a CUDA graph and replayed, which removes the launch gaps of the coarse levels
    struct GraphEntry {
        cudaGraphExec_t exec;
        TV* result;
        long long launches;
    };
    std::map<std::tuple<const TV*, TV*, TV*, bool, char>, GraphEntry> graphs;
    void invalidate_graphs() {
        for (auto& kv : graphs) cudaGraphExecDestroy(kv.second.exec);
        graphs.clear();
    }

    Hierarchy(int nlevels, int nrhs, char ct, int rk, const int64_t* rpre, const int64_t* rpost, int dev) {
        MGB_CHECK(nlevels >= 1, "levels must be >= 1");
        MGB_CHECK(nrhs >= 1, "nrhs must be >= 1");
        levels = nlevels;
        m = nrhs;
        relax_kind = rk;
        L.resize(levels);
        put_done.assign(levels, nullptr);
        set_cycle(ct, rpre, rpost);
        ctx.init(dev);
        use_fused_put = env_int("MGB200_FUSED_PUT", 1);
    }
    ~Hierarchy() override {
        cudaSetDevice(ctx.device);
        if (ctx.stream) cudaStreamSynchronize(ctx.stream);
        invalidate_graphs();
        p2p_release();
        for (auto& l : L) l.release();
        coarse.release();
        Akry.release();
        free_krylov();
        dev_free(hstage);
        dev_free(ux0);
        dev_free(ux1);
        dev_free(gram_ws);
        dev_free(gram_partials);
        dev_free(gram_counter);
        dev_free(kq);
        if (comm.comm) nccl().CommDestroy(comm.comm);
        comm.comm = nullptr;
        for (auto& e : mix_ev)
            if (e) cudaEventDestroy(e);
        ctx.destroy();
    }
    void set_cycle(char ct, const int64_t* rpre, const int64_t* rpost) {
        MGB_CHECK(ct == 'V' || ct == 'F' || ct == 'W' || ct == 'K', "cycle_type must be V, F, W or K");
        // the front ends set the cycle before every solve: unchanged parameters keep the workspaces and graphs
        if (ct == cycle_type && (int)pre.size() == levels && (int)post.size() == levels) {
            bool same = true;
            if (rpre && rpost)
                for (int l = 0; l < levels && same; ++l) same = (pre[l] == (int)rpre[l] && post[l] == (int)rpost[l]);
            if (same) return;
        }
        cycle_type = ct;
        invalidate_graphs();
        if (rpre && rpost) {
            pre.assign(levels, 0);
            post.assign(levels, 0);
            for (int l = 0; l < levels; ++l) {
                pre[l] = (int)rpre[l];
                post[l] = (int)rpost[l];
            }
        }
        work_ready = false;
    }
    void free_krylov() {
        dev_free(kr); dev_free(kp); dev_free(kAp); dev_free(kw); dev_free(kV); dev_free(kZ); dev_free(kq);
        kV_cols = 0;
    }

    // ---- upload ---------------------------------------------------------------------------
    void upload_level(int level, long long n, long long nc, const int64_t* acp, const int64_t* arv,
                      const void* anz, const int64_t* pcp, const int64_t* prv, const void* pnz,
                      const int64_t* rcp, const int64_t* rrv, const void* rnz, const void* d, int base) {
        MGB_CHECK(level >= 1 && level < levels, "upload_level: level must be in 1..levels-1");
        MGB_CHECK(base == 0 || base == 1, "index_base must be 0 or 1");
        MGB_CUDA(cudaSetDevice(ctx.device));
        Level<TV>& lv = L[level - 1];
        lv.n = n;
        upload_csr<TV>(ctx, lv.A, n, n, acp, arv, static_cast<const TV*>(anz), base, true);
        if (lv.ghint) {
            // the hint is only kept for the operator it describes exactly (grid_xfer.cuh); otherwise nothing changes
            HostPatterns<RT> hp;
            upload_csr<RT>(ctx, lv.P, n, nc, pcp, prv, static_cast<const RT*>(pnz), base, false, true, &hp);
            if (lv.P.pat.present && gx_verify_prolongation<RT>(hp, n, lv.gn, lv.gN, lv.P.gx)) gx_attach_table(lv.P, hp);
            hp = HostPatterns<RT>();
            upload_csr<RT>(ctx, lv.R, nc, n, rcp, rrv, static_cast<const RT*>(rnz), base, false, true, &hp);
            if (lv.R.pat.present && gx_verify_restriction<RT>(hp, nc, lv.gn, lv.gN, lv.R.gx)) gx_attach_table(lv.R, hp);
        } else {
            upload_csr<RT>(ctx, lv.P, n, nc, pcp, prv, static_cast<const RT*>(pnz), base, false);
            upload_csr<RT>(ctx, lv.R, nc, n, rcp, rrv, static_cast<const RT*>(rnz), base, false);
        }
        dev_free(lv.d);
        lv.d = dev_alloc<TV>(n + 4);   // slack for the even-rounded tile copies of the TMA kernel
        MGB_CUDA(cudaMemcpy(lv.d, d, n * sizeof(TV), cudaMemcpyHostToDevice));
        fold_d(lv, static_cast<const TV*>(d));
        lv.nalloc = n;
        lv.sp.dist = false;
        L[level].n = nc;
        L[level].nalloc = nc;
        work_ready = false;
    }

    // dense value table of a verified grid hint (grid_xfer.cuh)
    void gx_attach_table(Csr<RT>& M, const HostPatterns<RT>& hp) {
        if (!M.gx.ok) return;
        const std::vector<RT> t = gx_dense_table<RT>(M.gx, hp.val);
        RT* dt = dev_alloc<RT>(t.size());
        MGB_CUDA(cudaMemcpy(dt, t.data(), t.size() * sizeof(RT), cudaMemcpyHostToDevice));
        M.gx.tab = dt;
    }

    // the meshes of level l and l+1 (param.Meshes[l].n + 1 nodes per dimension), BEFORE upload_level(level)
    void set_level_grid(int level, int dim, const int64_t* nf, const int64_t* nc) {
        MGB_CHECK(level >= 1 && level < levels, "set_level_grid: level must be in 1..levels-1");
        MGB_CHECK(dim >= 1 && dim <= 3 && nf && nc, "set_level_grid: dim must be 1, 2 or 3");
        Level<TV>& lv = L[level - 1];
        for (int d = 0; d < 3; ++d) {
            lv.gn[d] = d < dim ? (int)nf[d] : 1;
            lv.gN[d] = d < dim ? (int)nc[d] : 1;
        }
        lv.ghint = true;
    }

    // ---- replaceMatrixInHierarchy on the device (galerkin.cuh; MGsetup.jl:226-270) -------------------------------------------
    // The new fine matrix (same sparsity as the resident one) comes from the host; relaxation weights, Galerkin products
    // and the coarsest factorisation are redone on the device with the resident Ps / Rs.  Returns false (nothing
    // changed) when the device path does not apply - another sparsity, a row-partitioned hierarchy, SPAI on a
    // structurally unsymmetric operator - and the caller redoes the setup on the host.
    // relax_type: 0 "Jac" / "Jac-GMRES", 1 "SPAI"; omega[l]: relaxParam of level l + 1.
    bool replace_matrix(long long n, const int64_t* cp, const int64_t* rv, const void* nz, int base, int relax_type,
                        const double* omega) {
        MGB_CHECK(levels >= 1 && n == L[0].n, "replace_matrix: size differs from the fine level");
        MGB_CHECK(base == 0 || base == 1, "index_base must be 0 or 1");
        MGB_CHECK(relax_type == 0 || relax_type == 1, "replace_matrix: relax_type must be 0 (Jac) or 1 (SPAI)");
        MGB_CUDA(cudaSetDevice(ctx.device));
        for (int l = 0; l < levels; ++l)
            if (L[l].sp.dist) return false;
        if (levels > 1 && !L[0].A.present()) return false;
        int* flag = dev_alloc<int>(1);
        auto read_flag = [&]() {
            int h = 0;
            MGB_CUDA(cudaMemcpyAsync(&h, flag, sizeof(int), cudaMemcpyDeviceToHost, ctx.stream));
            ctx.sync();
            MGB_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), ctx.stream));
            return h;
        };
        MGB_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), ctx.stream));
        // (1) the new fine matrix: a full upload (dictionary, box tables) next to the old one, kept only if the sparsity agrees
        Csr<TV> Anew;
        if (levels == 1) {
            upload_csr<TV>(ctx, L[0].A, n, n, cp, rv, static_cast<const TV*>(nz), base, true, false);
            factor_or_refresh_coarsest(omega ? omega[0] : 1.0);
            dev_free(flag);
            invalidate_graphs();
            return true;
        }
        upload_csr<TV>(ctx, Anew, n, n, cp, rv, static_cast<const TV*>(nz), base, true);
        Csr<TV>& A1 = L[0].A;
        bool same = Anew.nnz == A1.nnz;
        if (same) {
            csr_same_structure_kernel<<<ctx.ew_blocks(n + 1), 256, 0, ctx.stream>>>(n + 1, Anew.rowptr, A1.rowptr, flag);
            csr_same_structure_kernel<<<ctx.ew_blocks(A1.nnz), 256, 0, ctx.stream>>>(A1.nnz, Anew.colind, A1.colind, flag);
            MGB_LAUNCH_CHECK();
            same = read_flag() == 0;
        }
        if (!same) {
            Anew.release();
            dev_free(flag);
            return false;
        }
        // SPAI needs the columns of A through a symmetric structure: test on every level before anything is changed
        std::vector<TV*> dnew(levels - 1, nullptr);
        auto fail = [&]() {
            for (auto& q : dnew) dev_free(q);
            Anew.release();
            dev_free(flag);
            return false;
        };
        // (2) level by level: weights of A_l, then A_{l+1} = R_l A_l P_l into the resident structure
        std::swap(A1, Anew);                  // A1 is the new matrix from here on; Anew holds the old one
        for (int l = 0; l + 1 < levels; ++l) {
            Level<TV>& lv = L[l];
            Csr<TV>& A = lv.A;
            dnew[l] = dev_alloc<TV>(lv.n + 4);
            MGB_CUDA(cudaMemsetAsync(dnew[l] + lv.n, 0, 4 * sizeof(TV), ctx.stream));
            relax_prec_kernel<TV><<<ctx.ew_blocks(lv.n), 256, 0, ctx.stream>>>((int)lv.n, A.rowptr, A.colind, A.val, relax_type,
                                                                                omega ? omega[l] : 1.0, dnew[l], flag);
            MGB_LAUNCH_CHECK();
            if (read_flag() != 0) {
                // restore: the old fine matrix goes back, coarse values already overwritten are recomputed from it
                std::swap(A1, Anew);
                for (int q = 0; q < l; ++q) galerkin_level(q);
                return fail();
            }
            galerkin_level(l);
        }
        Anew.release();
        // (3) commit: weights, dictionaries of the coarse operators, coarsest factorisation
        for (int l = 0; l + 1 < levels; ++l) {
            Level<TV>& lv = L[l];
            dev_free(lv.d);
            lv.d = dnew[l];
            if (l > 0) refresh_dictionary(lv.A, flag);
            fold_d_device(lv, flag);
        }
        factor_or_refresh_coarsest(omega ? omega[levels - 1] : 1.0);
        dev_free(flag);
        invalidate_graphs();
        return true;
    }
    void galerkin_level(int l) {
        Level<TV>& lv = L[l];
        Csr<TV>& C = L[l + 1].A;
        MGB_CHECK(C.present() && lv.R.present() && lv.P.present(), "replace_matrix: level not uploaded");
        const int nc = C.n_rows;
        const int grid = std::min(cdiv(nc, 8), ctx.sm_count * 8 * 4);
        galerkin_kernel<TV, RT><<<grid, 256, 0, ctx.stream>>>(nc, lv.R.rowptr, lv.R.colind, lv.R.val, lv.A.rowptr, lv.A.colind,
                                                               lv.A.val, lv.P.rowptr, lv.P.colind, lv.P.val, C.rowptr, C.colind, C.val);
        MGB_LAUNCH_CHECK();
    }
    // dictionary / box tables of a matrix whose CSR values changed: kept when every row still equals its pattern's
    // representative row, dropped (CSR-stream kernels from then on) otherwise
    void refresh_dictionary(Csr<TV>& M, int* flag) {
        PatDict<TV>& D = M.pat;
        if (!D.present) return;
        bool ok = D.rep != nullptr;
        if (ok) {
            pat_verify_values_kernel<TV><<<ctx.ew_blocks(M.n_rows), 256, 0, ctx.stream>>>(M.n_rows, M.rowptr, M.val, D.pid, D.rep, flag);
            MGB_LAUNCH_CHECK();
            int h = 0;
            MGB_CUDA(cudaMemcpyAsync(&h, flag, sizeof(int), cudaMemcpyDeviceToHost, ctx.stream));
            ctx.sync();
            MGB_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), ctx.stream));
            ok = h == 0;
        }
        if (!ok) {
            D.release();
            M.box.release();
            return;
        }
        TV* vals = dev_alloc<TV>(std::max(D.nent, 1));
        pat_refresh_values_kernel<TV><<<D.npat, 32, 0, ctx.stream>>>(D.npat, D.pat_off, D.rep, M.rowptr, M.val, D.ent, D.ent_s, vals);
        MGB_LAUNCH_CHECK();
        BoxDict<TV>& X = M.box;
        if (X.ok) {
            std::vector<TV> hv(std::max(D.nent, 1));
            MGB_CUDA(cudaMemcpyAsync(hv.data(), vals, (size_t)D.nent * sizeof(TV), cudaMemcpyDeviceToHost, ctx.stream));
            ctx.sync();
            if (!X.refill(hv)) {
                X.release();
            } else {
                MGB_CUDA(cudaMemcpy(X.ctab, X.h_ctab.data(), X.h_ctab.size() * sizeof(TV), cudaMemcpyHostToDevice));
            }
        }
        ctx.sync();
        dev_free(vals);
    }
    // fold_d from the resident d (no host copy): d[rep[p]] per pattern, verified against every row
    void fold_d_device(Level<TV>& lv, int* flag) {
        dev_free(lv.dpat);
        PatDict<TV>& D = lv.A.pat;
        D.host_pid.clear();
        D.host_pid.shrink_to_fit();
        if (!D.present || !D.rep || env_int("MGB200_FOLD_D", 1) == 0) return;
        TV* dp = dev_alloc<TV>(D.npat);
        pat_gather_d_kernel<TV><<<cdiv(D.npat, 64), 64, 0, ctx.stream>>>(D.npat, D.rep, lv.d, dp);
        pat_verify_d_kernel<TV><<<ctx.ew_blocks(lv.n), 256, 0, ctx.stream>>>((int)lv.n, D.pid, dp, lv.d, flag);
        MGB_LAUNCH_CHECK();
        int h = 0;
        MGB_CUDA(cudaMemcpyAsync(&h, flag, sizeof(int), cudaMemcpyDeviceToHost, ctx.stream));
        ctx.sync();
        MGB_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), ctx.stream));
        if (h != 0) {
            dev_free(dp);
            return;
        }
        lv.dpat = dp;
        BoxDict<TV>& X = lv.A.box;
        if (X.ok) {
            std::vector<TV> hd(D.npat);
            MGB_CUDA(cudaMemcpy(hd.data(), dp, D.npat * sizeof(TV), cudaMemcpyDeviceToHost));
            MGB_CUDA(cudaMemcpy(X.dtab, hd.data(), D.npat * sizeof(TV), cudaMemcpyHostToDevice));
            X.c0.d0 = hd[X.p0];
        }
    }
    // coarsest solver of the resident As[end]: dense LU again, or (coarseSolveType "GMRES") the Jacobi weights
    void factor_or_refresh_coarsest(double omega) {
        if (coarse.kind == 1) {
            Level<TV>& lv = L[levels - 1];
            int* flag = dev_alloc<int>(1);
            MGB_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), ctx.stream));
            relax_prec_kernel<TV><<<ctx.ew_blocks(lv.n), 256, 0, ctx.stream>>>((int)lv.n, lv.A.rowptr, lv.A.colind, lv.A.val, 0, omega,
                                                                                coarse.d, flag);
            MGB_LAUNCH_CHECK();
            ctx.sync();
            dev_free(flag);
            refresh_dictionary(lv.A, nullptr_flag());
        } else {
            factor_coarsest();
        }
    }
    int* scratch_flag = nullptr;
    int* nullptr_flag() {
        if (!scratch_flag) {
            scratch_flag = dev_alloc<int>(1);
            MGB_CUDA(cudaMemset(scratch_flag, 0, sizeof(int)));
        }
        return scratch_flag;
    }
    // values of a resident matrix back in the caller's convention (nzval of the stored adjoint: conjugated for A)
    void download_values(int level, int which, void* out, long long nnz) {
        MGB_CHECK(level >= 1 && level <= levels, "download_values: level out of range");
        MGB_CUDA(cudaSetDevice(ctx.device));
        Level<TV>& lv = L[level - 1];
        ctx.sync();
        if (which == 0) {
            MGB_CHECK(lv.A.present() && lv.A.nnz == nnz, "download_values: nnz differs from the resident matrix");
            std::vector<TV> h((size_t)nnz);
            MGB_CUDA(cudaMemcpy(h.data(), lv.A.val, (size_t)nnz * sizeof(TV), cudaMemcpyDeviceToHost));
            TV* o = static_cast<TV*>(out);
            for (long long k = 0; k < nnz; ++k) o[k] = conj_(h[k]);
        } else {
            Csr<RT>& M = which == 1 ? lv.P : lv.R;
            MGB_CHECK(M.present() && M.nnz == nnz, "download_values: nnz differs from the resident matrix");
            MGB_CUDA(cudaMemcpy(out, M.val, (size_t)nnz * sizeof(RT), cudaMemcpyDeviceToHost));
        }
    }
    void download_relax_prec(int level, void* out) {
        MGB_CHECK(level >= 1 && level < levels, "download_relax_prec: level must be in 1..levels-1");
        MGB_CUDA(cudaSetDevice(ctx.device));
        ctx.sync();
        MGB_CHECK(L[level - 1].d, "download_relax_prec: level not uploaded");
        MGB_CUDA(cudaMemcpy(out, L[level - 1].d, (size_t)L[level - 1].n * sizeof(TV), cudaMemcpyDeviceToHost));
    }

    // relaxPrecs[l] as a function of A_l's pattern id: d folds into the dictionary when every row of
    // a pattern carries bit-identical d (constant-coefficient stencils)
    void fold_d(Level<TV>& lv, const TV* d) {
        dev_free(lv.dpat);
        PatDict<TV>& D = lv.A.pat;
        if (!D.present || D.host_pid.empty()) return;
        if (env_int("MGB200_FOLD_D", 1) == 0) {   // tests: keep d as a vector
            D.host_pid.clear();
            D.host_pid.shrink_to_fit();
            return;
        }
        std::vector<TV> dp(D.npat);
        std::vector<char> seen(D.npat, 0);
        bool ok = true;
        for (long long i = 0; i < lv.n && ok; ++i) {
            const int p = D.host_pid[i];
            if (!seen[p]) {
                seen[p] = 1;
                dp[p] = d[i];
            } else if (std::memcmp(&dp[p], &d[i], sizeof(TV)) != 0) {
                ok = false;
            }
        }
        D.host_pid.clear();
        D.host_pid.shrink_to_fit();
        if (!ok) return;
        lv.dpat = dev_alloc<TV>(D.npat);
        MGB_CUDA(cudaMemcpy(lv.dpat, dp.data(), D.npat * sizeof(TV), cudaMemcpyHostToDevice));
        BoxDict<TV>& X = lv.A.box;
        if (X.ok) {      // the box-stencil kernel keeps its own (padded) copy and the dominant pattern's weight
            MGB_CUDA(cudaMemcpy(X.dtab, dp.data(), D.npat * sizeof(TV), cudaMemcpyHostToDevice));
            X.c0.d0 = dp[X.p0];
        }
    }

    void upload_coarsest(long long n, const int64_t* cp, const int64_t* rv, const void* nz, int base) {
        MGB_CHECK(n >= 1 && n <= 46000, "coarsest grid size out of range for the dense device LU");
        MGB_CUDA(cudaSetDevice(ctx.device));
        Level<TV>& lv = L[levels - 1];
        lv.n = n;
        lv.nalloc = n;
        upload_csr<TV>(ctx, lv.A, n, n, cp, rv, static_cast<const TV*>(nz), base, true, false);
        factor_coarsest();
    }
    // dense LU of the resident As[end] (also after replace_matrix put new values into it)
    void factor_coarsest() {
        Level<TV>& lv = L[levels - 1];
        const long long n = lv.n;
        coarse.release();
        coarse.n = (int)n;
        const int N = (int)n;
        typedef typename Wide<TV>::type TW;
        TW* a = dev_alloc<TW>((size_t)N * N);
        MGB_CUDA(cudaMemsetAsync(a, 0, (size_t)N * N * sizeof(TW), ctx.stream));
        densify_kernel<TV, TW><<<cdiv(N, 128), 128, 0, ctx.stream>>>(N, lv.A.rowptr, lv.A.colind, lv.A.val, a);
        MGB_LAUNCH_CHECK();
        int* piv = dev_alloc<int>(N + 1);
        int* info = piv + N;
        MGB_CUDA(cudaMemsetAsync(info, 0, sizeof(int), ctx.stream));
        // blocked right-looking LU (dense_lu.cuh): per panel of LU_NB columns one cooperative panel kernel, the row
        // interchanges outside the panel, the triangular solve for U12 and the trailing update as a tiled product
        {
            int per_sm = 0;
            MGB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lu_panel_kernel<TW>, 256, 0));
            MGB_CHECK(per_sm >= 1, "LU panel kernel does not fit an SM");
            const int pg = std::max(1, std::min(ctx.sm_count * std::min(per_sm, 2), cdiv(N, 256)));
            LuPivot* cand = dev_alloc<LuPivot>(pg);
            unsigned* counter = dev_alloc<unsigned>(1);
            MGB_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned), ctx.stream));
            unsigned epoch = 0;
            for (int k0 = 0; k0 < N; k0 += LU_NB) {
                int nb = std::min(LU_NB, N - k0);
                TW* ap = a;
                int nn = N, kk = k0;
                void* args[] = {&ap, &nn, &kk, &nb, &piv, &info, &cand, &counter, &epoch};
                MGB_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(lu_panel_kernel<TW>), dim3(pg), dim3(256), args, 0,
                                                     ctx.stream));
                epoch += 3u * (unsigned)nb * (unsigned)pg;          // three grid barriers per column
                if (N > nb) lu_swap_kernel<TW><<<cdiv(N - nb, 256), 256, 0, ctx.stream>>>(a, N, k0, nb, piv);
                const int rem = N - k0 - nb;
                if (rem > 0) {
                    lu_trsm_kernel<TW><<<cdiv(rem, 128), 128, 0, ctx.stream>>>(a, N, k0, nb);
                    dim3 grd(cdiv(rem, LuTile<TW>::T), cdiv(rem, LuTile<TW>::T));
                    lu_gemm_kernel<TW><<<grd, 256, 0, ctx.stream>>>(a, N, k0, nb);
                }
            }
            MGB_LAUNCH_CHECK();
            ctx.sync();
            dev_free(cand);
            dev_free(counter);
        }
        std::vector<int> hpiv(N + 1);
        MGB_CUDA(cudaMemcpyAsync(hpiv.data(), piv, (N + 1) * sizeof(int), cudaMemcpyDeviceToHost, ctx.stream));
        ctx.sync();
        MGB_CHECK(hpiv[N] == 0, "coarsest matrix is exactly singular (zero pivot in LU)");
        std::vector<int> perm(N);
        for (int i = 0; i < N; ++i) perm[i] = i;
        for (int k = 0; k < N; ++k) std::swap(perm[k], perm[hpiv[k]]);
        coarse.perm = dev_alloc<int>(N);
        MGB_CUDA(cudaMemcpy(coarse.perm, perm.data(), N * sizeof(int), cudaMemcpyHostToDevice));
        coarse.linv = dev_alloc<TW>((size_t)N * N);
        coarse.uinv = dev_alloc<TW>((size_t)N * N);
        MGB_CUDA(cudaMemsetAsync(coarse.linv, 0, (size_t)N * N * sizeof(TW), ctx.stream));
        MGB_CUDA(cudaMemsetAsync(coarse.uinv, 0, (size_t)N * N * sizeof(TW), ctx.stream));
        // blocked inversion of the two factors: block rows in dependency order (L top-down, U bottom-up)
        const int nblk = cdiv(N, TI_NB);
        for (int bi = 0; bi < nblk; ++bi) {
            const int r0 = bi * TI_NB, nbk = std::min(TI_NB, N - r0);
            tri_diag_inv_kernel<TW, false><<<1, TI_NB, 0, ctx.stream>>>(a, N, r0, nbk, coarse.linv);
            if (r0 > 0) tri_inv_row_kernel<TW, false><<<cdiv(r0, LuTile<TW>::T), 256, 0, ctx.stream>>>(a, N, r0, nbk, coarse.linv);
        }
        for (int bi = nblk - 1; bi >= 0; --bi) {
            const int r0 = bi * TI_NB, nbk = std::min(TI_NB, N - r0);
            tri_diag_inv_kernel<TW, true><<<1, TI_NB, 0, ctx.stream>>>(a, N, r0, nbk, coarse.uinv);
            const int right = N - r0 - nbk;
            if (right > 0) tri_inv_row_kernel<TW, true><<<cdiv(right, LuTile<TW>::T), 256, 0, ctx.stream>>>(a, N, r0, nbk, coarse.uinv);
        }
        MGB_LAUNCH_CHECK();
        ctx.sync();
        dev_free(a);
        dev_free(piv);
        work_ready = false;
    }

    // defineCoarsestAinv, coarseSolveType == "GMRES" (MGsetup.jl:333-334): As[end] and the Jacobi weights d
    void upload_coarsest_gmres(long long n, const int64_t* cp, const int64_t* rv, const void* nz, const void* d,
                               int base) {
        MGB_CHECK(n >= 2, "coarsest grid too small for GMRES");
        MGB_CUDA(cudaSetDevice(ctx.device));
        Level<TV>& lv = L[levels - 1];
        lv.n = n;
        lv.nalloc = n;
        upload_csr<TV>(ctx, lv.A, n, n, cp, rv, static_cast<const TV*>(nz), base, true);
        coarse.release();
        coarse.n = (int)n;
        coarse.kind = 1;
        coarse.d = dev_alloc<TV>(n);
        MGB_CUDA(cudaMemcpy(coarse.d, d, n * sizeof(TV), cudaMemcpyHostToDevice));
        const int restrt = (int)std::min<long long>(10, n - 1);
        coarse.r = dev_alloc<TV>(n);
        coarse.w = dev_alloc<TV>(n);
        coarse.t = dev_alloc<TV>(n);
        coarse.dv = dev_alloc<TV>(n);
        coarse.V = dev_alloc<TV>((size_t)n * restrt);
        invalidate_graphs();
        work_ready = false;
    }

    void set_krylov_matrix(long long n, const int64_t* cp, const int64_t* rv, const void* nz, int base) {
        MGB_CUDA(cudaSetDevice(ctx.device));
        if (cp == nullptr || n == 0) {   // back to the hierarchy's own fine matrix (getAfun(AT) with AT === As[1])
            Akry.release();
            return;
        }
        MGB_CHECK(n == L[0].n, "Krylov matrix size differs from the fine level");
        upload_csr<TV>(ctx, Akry, n, n, cp, rv, static_cast<const TV*>(nz), base, true);
    }

    // ---- multi-GPU: communicator, staged slab upload, finalisation -----------------------------
    void dist_init(int rank, int world, const char* unique_id) {
        MGB_CHECK(world >= 1 && rank >= 0 && rank < world, "bad rank / world");
        MGB_CUDA(cudaSetDevice(ctx.device));
        comm.rank = rank;
        comm.world = world;
        if (world > 1) {
            nccl().load();
            ncclUniqueId id;
            std::memcpy(&id, unique_id, sizeof(id));
            MGB_NCCL(nccl().CommInitRank(&comm.comm, world, id, rank));
        }
    }

    // stage the owned rows of level `level` (1-based): CSC-of-adjoint blocks whose columns are the
    // owned rows and whose row indices are GLOBAL operator-column indices.
    template <typename TA>
    static void stage_rows(HostRows<TA>& H, long long n_rows, const int64_t* cp, const int64_t* rv,
                           const TA* nz, int base, bool conjugate) {
        H.n_rows = n_rows;
        H.rowptr.resize(n_rows + 1);
        for (long long i = 0; i <= n_rows; ++i) H.rowptr[i] = cp[i] - base;
        const long long nnz = H.rowptr[n_rows];
        H.col.resize(nnz);
        H.val.resize(nnz);
        for (long long k = 0; k < nnz; ++k) {
            H.col[k] = rv[k] - base;
            H.val[k] = conjugate ? conj_(nz[k]) : nz[k];
        }
    }
    void dist_upload_level(int level, long long n_global, const int64_t* row_offsets, long long nc_global,
                           const int64_t* coarse_row_offsets, const int64_t* acp, const int64_t* arv,
                           const void* anz, const int64_t* pcp, const int64_t* prv, const void* pnz,
                           const int64_t* rcp, const int64_t* rrv, const void* rnz, const void* d, int base) {
        MGB_CHECK(level >= 1 && level < levels, "dist_upload_level: level must be in 1..levels-1");
        MGB_CHECK(base == 0 || base == 1, "index_base must be 0 or 1");
        Level<TV>& lv = L[level - 1];
        const int w = comm.world, r = comm.rank;
        lv.sp.dist = true;
        lv.sp.n_global = n_global;
        lv.sp.row_offsets.assign(row_offsets, row_offsets + w + 1);
        lv.sp.lo = row_offsets[r];
        lv.sp.hi = row_offsets[r + 1];
        lv.sp.n_owned = lv.sp.hi - lv.sp.lo;
        lv.n = lv.sp.n_owned;
        lv.nc_global = nc_global;
        lv.coarse_row_offsets.assign(coarse_row_offsets, coarse_row_offsets + w + 1);
        const long long nc_owned = coarse_row_offsets[r + 1] - coarse_row_offsets[r];
        stage_rows<TV>(lv.hA, lv.n, acp, arv, static_cast<const TV*>(anz), base, true);
        stage_rows<RT>(lv.hP, lv.n, pcp, prv, static_cast<const RT*>(pnz), base, false);
        stage_rows<RT>(lv.hR, nc_owned, rcp, rrv, static_cast<const RT*>(rnz), base, false);
        lv.hd.assign(static_cast<const TV*>(d), static_cast<const TV*>(d) + lv.n);
        dist_finalized = false;
        work_ready = false;
    }

    // Build the ghost sets, exchange the request lists, remap the columns into the
    // [owned | ghost] layouts and upload the CSR matrices.  Collective over all ranks.
    void finalize_dist() {
        const int w = comm.world, r = comm.rank;
        // 1. ghost union per distributed level space: columns of A_l, R_l and P_{l-1}
        for (int l = 0; l < levels - 1; ++l) {
            Level<TV>& lv = L[l];
            if (!lv.sp.dist) continue;
            MGB_CHECK(lv.hA.present(), "distributed level staged twice or not at all");
            DistSpace& sp = lv.sp;
            sp.ghosts.clear();
            collect_ghosts(lv.hA.col.data(), (long long)lv.hA.col.size(), sp.lo, sp.hi, sp.ghosts);
            collect_ghosts(lv.hR.col.data(), (long long)lv.hR.col.size(), sp.lo, sp.hi, sp.ghosts);
            if (l > 0 && L[l - 1].sp.dist)
                collect_ghosts(L[l - 1].hP.col.data(), (long long)L[l - 1].hP.col.size(), sp.lo, sp.hi, sp.ghosts);
            sort_unique(sp.ghosts);
            sp.n_ghost = (long long)sp.ghosts.size();
            sp.n_lo = std::lower_bound(sp.ghosts.begin(), sp.ghosts.end(), sp.lo) - sp.ghosts.begin();
            lv.nalloc = sp.n_owned + sp.n_ghost;
            MGB_CHECK(lv.nalloc < (1LL << 31) - 8, "local vector too long for 32-bit indices");
            sp.recv_cnt.assign(w, 0);
            sp.recv_off.assign(w, 0);
            for (long long g : sp.ghosts) {
                MGB_CHECK(g >= 0 && g < sp.n_global, "column index outside the global range");
                sp.recv_cnt[sp.owner_of(g)]++;
            }
            for (int p = 1; p < w; ++p) sp.recv_off[p] = sp.recv_off[p - 1] + sp.recv_cnt[p - 1];
            MGB_CHECK(sp.recv_cnt[r] == 0, "ghost owned by self");
        }
        // 2. request exchange (who needs what from me) over NCCL
        for (int l = 0; l < levels - 1; ++l) {
            Level<TV>& lv = L[l];
            if (!lv.sp.dist) continue;
            DistSpace& sp = lv.sp;
            sp.send_cnt.assign(w, 0);
            sp.send_off.assign(w, 0);
            sp.release();
            if (w == 1) continue;
            int* d_cnt = dev_alloc<int>((size_t)w * w);
            MGB_CUDA(cudaMemcpyAsync(d_cnt + (size_t)r * w, sp.recv_cnt.data(), w * sizeof(int), cudaMemcpyHostToDevice, ctx.stream));
            MGB_NCCL(nccl().AllGather(d_cnt + (size_t)r * w, d_cnt, (size_t)w, ncclInt32, comm.comm, ctx.stream));
            std::vector<int> all((size_t)w * w);
            MGB_CUDA(cudaMemcpyAsync(all.data(), d_cnt, (size_t)w * w * sizeof(int), cudaMemcpyDeviceToHost, ctx.stream));
            ctx.sync();
            dev_free(d_cnt);
            for (int q = 0; q < w; ++q) sp.send_cnt[q] = all[(size_t)q * w + r];
            for (int q = 1; q < w; ++q) sp.send_off[q] = sp.send_off[q - 1] + sp.send_cnt[q - 1];
            sp.n_send = sp.send_off[w - 1] + sp.send_cnt[w - 1];
            long long* d_req = dev_alloc<long long>(std::max<long long>(sp.n_ghost, 1));
            long long* d_got = dev_alloc<long long>(std::max(sp.n_send, 1));
            MGB_CUDA(cudaMemcpyAsync(d_req, sp.ghosts.data(), sp.n_ghost * sizeof(long long), cudaMemcpyHostToDevice, ctx.stream));
            MGB_NCCL(nccl().GroupStart());
            for (int p = 0; p < w; ++p) {
                if (sp.recv_cnt[p] > 0) MGB_NCCL(nccl().Send(d_req + sp.recv_off[p], sp.recv_cnt[p], ncclInt64, p, comm.comm, ctx.stream));
                if (sp.send_cnt[p] > 0) MGB_NCCL(nccl().Recv(d_got + sp.send_off[p], sp.send_cnt[p], ncclInt64, p, comm.comm, ctx.stream));
            }
            MGB_NCCL(nccl().GroupEnd());
            std::vector<long long> got(std::max(sp.n_send, 1));
            MGB_CUDA(cudaMemcpyAsync(got.data(), d_got, sp.n_send * sizeof(long long), cudaMemcpyDeviceToHost, ctx.stream));
            ctx.sync();
            dev_free(d_req);
            dev_free(d_got);
            std::vector<int> idx(std::max(sp.n_send, 1));
            for (int k = 0; k < sp.n_send; ++k) {
                MGB_CHECK(got[k] >= sp.lo && got[k] < sp.hi, "peer requested a row this rank does not own");
                idx[k] = (int)(got[k] - sp.lo);
            }
            sp.d_send_idx = dev_alloc<int>(std::max(sp.n_send, 1));
            MGB_CUDA(cudaMemcpy(sp.d_send_idx, idx.data(), sp.n_send * sizeof(int), cudaMemcpyHostToDevice));
            sp.send_idx_host.assign(idx.begin(), idx.begin() + sp.n_send);
        }
        // 3. remap columns and upload
        for (int l = 0; l < levels - 1; ++l) {
            Level<TV>& lv = L[l];
            if (!lv.sp.dist) continue;
            DistSpace& sp = lv.sp;
            for (auto& c : lv.hA.col) c = to_local_split(c, sp.lo, sp.hi, sp.ghosts, sp.n_lo);
            for (auto& c : lv.hR.col) c = to_local_split(c, sp.lo, sp.hi, sp.ghosts, sp.n_lo);
            Level<TV>& lc = L[l + 1];
            long long pcols;
            if (lc.sp.dist) {
                // lc's ghost set already contains P's columns (step 1)
                for (auto& c : lv.hP.col) c = to_local_split(c, lc.sp.lo, lc.sp.hi, lc.sp.ghosts, lc.sp.n_lo);
                pcols = lc.sp.n_owned + lc.sp.n_ghost;
            } else {
                pcols = lv.nc_global;
                MGB_CHECK(lc.n == 0 || lc.n == lv.nc_global, "replicated coarse level size mismatch");
                lc.n = lv.nc_global;
                if (lc.nalloc < lc.n) lc.nalloc = lc.n;
            }
            upload_csr<TV>(ctx, lv.A, lv.hA.n_rows, lv.nalloc, lv.hA.rowptr.data(), lv.hA.col.data(), lv.hA.val.data(), 0, false);
            if (lv.ghint && lv.gn[2] > 1) {
                // grid hint on a z-slab (grid_xfer.cuh): the rows are whole planes of the global grids, the local column
                // indices are the global ones minus the first owned row of the column space - checked row by row
                const long long pf = (long long)lv.gn[0] * lv.gn[1], pc = (long long)lv.gN[0] * lv.gN[1];
                const long long clo = lv.coarse_row_offsets[comm.rank], chi = lv.coarse_row_offsets[comm.rank + 1];
                HostPatterns<RT> hp;
                upload_csr<RT>(ctx, lv.P, lv.hP.n_rows, pcols, lv.hP.rowptr.data(), lv.hP.col.data(), lv.hP.val.data(), 0, false, true, &hp);
                if (lv.P.pat.present && sp.lo % pf == 0 && sp.n_owned % pf == 0 &&
                    gx_verify_prolongation<RT>(hp, lv.hP.n_rows, lv.gn, lv.gN, lv.P.gx, (int)(sp.lo / pf), (int)(sp.n_owned / pf),
                                               lc.sp.dist ? lc.sp.lo : 0))
                    gx_attach_table(lv.P, hp);
                hp = HostPatterns<RT>();
                upload_csr<RT>(ctx, lv.R, lv.hR.n_rows, lv.nalloc, lv.hR.rowptr.data(), lv.hR.col.data(), lv.hR.val.data(), 0, false, true, &hp);
                if (lv.R.pat.present && clo % pc == 0 && (chi - clo) % pc == 0 &&
                    gx_verify_restriction<RT>(hp, lv.hR.n_rows, lv.gn, lv.gN, lv.R.gx, (int)(clo / pc), (int)((chi - clo) / pc), sp.lo))
                    gx_attach_table(lv.R, hp);
            } else {
                upload_csr<RT>(ctx, lv.P, lv.hP.n_rows, pcols, lv.hP.rowptr.data(), lv.hP.col.data(), lv.hP.val.data(), 0, false);
                upload_csr<RT>(ctx, lv.R, lv.hR.n_rows, lv.nalloc, lv.hR.rowptr.data(), lv.hR.col.data(), lv.hR.val.data(), 0, false);
            }
            // rows that read no ghost row of their input vector (they may run beside the halo exchange, apply_x)
            interior_rows(lv.hA, sp.n_owned, lv.A.int_lo, lv.A.int_hi);
            interior_rows(lv.hR, sp.n_owned, lv.R.int_lo, lv.R.int_hi);
            if (lc.sp.dist) interior_rows(lv.hP, lc.sp.n_owned, lv.P.int_lo, lv.P.int_hi);
            dev_free(lv.d);
            lv.d = dev_alloc<TV>(lv.n + 4);
            MGB_CUDA(cudaMemcpy(lv.d, lv.hd.data(), lv.n * sizeof(TV), cudaMemcpyHostToDevice));
            fold_d(lv, lv.hd.data());
            // input vectors of A_l are laid out [ghosts below | owned | ghosts above] around the vector pointer
            lv.A.pat.xlo = -(long long)align_pad<TV>((size_t)sp.n_lo);
            lv.A.pat.xhi = (long long)align_pad<TV>((size_t)(sp.n_owned + (sp.n_ghost - sp.n_lo)));
        }
        for (int l = 0; l < levels - 1; ++l) {
            L[l].hA.clear();
            L[l].hP.clear();
            L[l].hR.clear();
            L[l].hd.clear();
            L[l].hd.shrink_to_fit();
        }
        dist_finalized = true;
    }

    // ---- workspaces (adjustMemoryForNumRHS, MGsetup.jl:166-223) -----------------------------
    void adjust_nrhs(int nrhs) {
        MGB_CHECK(nrhs >= 1, "nrhs must be >= 1");
        if (nrhs != m) {
            m = nrhs;
            work_ready = false;
        }
    }
    void ensure_work() {
        if (work_ready) return;
        MGB_CUDA(cudaSetDevice(ctx.device));
        invalidate_graphs();
        if (!dist_finalized) finalize_dist();
        for (int l = 0; l < levels; ++l) {
            Level<TV>& lv = L[l];
            MGB_CHECK(lv.n > 0, "hierarchy level not uploaded");
            if (l < levels - 1) MGB_CHECK(lv.A.present() && lv.P.present() && lv.R.present(), "hierarchy level not uploaded");
            lv.release_work();
            if (lv.nalloc < lv.n) lv.nalloc = lv.n;
            const size_t nm = (size_t)lv.nalloc * m;
            if (lv.sp.dist && lv.sp.n_send > 0) {
                if (lv.sp.sendbuf) cudaFree(lv.sp.sendbuf);
                lv.sp.sendbuf_bytes = (size_t)lv.sp.n_send * m * sizeof(TV);
                MGB_CUDA(cudaMalloc(&lv.sp.sendbuf, lv.sp.sendbuf_bytes));
            }
            const size_t pad = vec_pad(l);
            lv.b = vec_alloc<TV>(nm, pad, ctx.stream);
            lv.r = vec_alloc<TV>(nm, pad, ctx.stream);
            lv.x0 = vec_alloc<TV>(nm, pad, ctx.stream);
            lv.x1 = vec_alloc<TV>(nm, pad, ctx.stream);
            if (l < levels - 1 && relax_kind == 1) {
                alloc_fgmres(lv.memRelax, nm, std::max(std::max(pre[l], post[l]), 1), pad);
            }
            // memKcycle[level] of the reference belongs to level+1 (MGsetup.jl:213-215): sized here
            // for level l (1-based l+1 >= 2) and used when the parent recurses into it.
            if (cycle_type == 'K' && l >= 1 && l < levels - 1) alloc_fgmres(lv.memK, nm, 2, pad);
        }
        MGB_CHECK(krylov_only || coarse.n == (int)L[levels - 1].n, "coarsest factorisation missing (mgb200_upload_coarsest)");
        MGB_CHECK(!krylov_only || Akry.present(), "mixed-precision handle without its Krylov matrix (mgb200_set_krylov_matrix)");
        dev_free(coarse.y);
        coarse.y = dev_alloc<typename Wide<TV>::type>((size_t)coarse.n * m);
        free_krylov();
        dev_free(hstage);
        hstage_n = 0;
        dev_free(ux0);
        dev_free(ux1);
        ux0 = vec_alloc<TV>((size_t)L[0].nalloc * m, vec_pad(0), ctx.stream);
        ux1 = vec_alloc<TV>((size_t)L[0].nalloc * m, vec_pad(0), ctx.stream);
        ucur = ux0;
        ctx.sync();
        p2p_setup();
        work_ready = true;
    }
    // elements in front of the first owned row of a level-l vector (lower ghost rows, dist.cuh)
    size_t vec_pad(int l) const { return L[l].sp.dist ? align_pad<TV>((size_t)L[l].sp.n_lo * m) : 0; }
    void alloc_fgmres(FgmresMem<TV>& mem, size_t nm, int inner, size_t pad) {
        mem.release();
        mem.inner = inner;
        mem.Z = vec_alloc<TV>(nm * inner, pad, ctx.stream);
        mem.AZ = vec_alloc<TV>(nm * inner, pad, ctx.stream);
        mem.Az = vec_alloc<TV>(nm, pad, ctx.stream);
        mem.vp0 = vec_alloc<TV>(nm, pad, ctx.stream);
        mem.vp1 = vec_alloc<TV>(nm, pad, ctx.stream);
    }
    void ensure_krylov(int vcols, bool needZ) {
        const size_t nm = (size_t)L[0].nalloc * m;
        const size_t pad = vec_pad(0);
        if (!kr) {
            kr = vec_alloc<TV>(nm, pad, ctx.stream);
            kp = vec_alloc<TV>(nm, pad, ctx.stream);
            kAp = vec_alloc<TV>(nm, pad, ctx.stream);
            kw = vec_alloc<TV>(nm, pad, ctx.stream);
        }
        if (vcols > kV_cols || (needZ && !kZ)) {
            dev_free(kV);
            dev_free(kZ);
            kV_cols = std::max(vcols, kV_cols);
            kV = vec_alloc<TV>(nm * kV_cols, pad, ctx.stream);
            kZ = vec_alloc<TV>(nm * kV_cols, pad, ctx.stream);
        }
    }

    // ---- building blocks ----------------------------------------------------------------------
    const Csr<TV>& krylov_A() const { return Akry.present() ? Akry : L[0].A; }

    void apply_A(const Csr<TV>& A, const TV* x, TV* y, int level) {  // y = A x   (getAfun, SolveFuncs.jl:65-71)
        apply_x<TV>(level - 1, A, MODE_SPMV, const_cast<TV*>(x), nullptr, nullptr, y, K_SPMV, level);
    }
    // r = b - A x; put_level >= 0: r is a level-(put_level+1) vector that is exchanged next (fused put)
    void residual(const Csr<TV>& A, const TV* b, const TV* x, TV* r, int level, int put_level = -1) {
        apply_x<TV>(level - 1, A, MODE_RESID, const_cast<TV*>(x), b, nullptr, r, K_RESID, level, nullptr, put_level);
    }
    // Plan for a kernel that produces level-l vector `out` which the NEXT operation on channel l exchanges: the
    // kernel stores the slab-end rows to the neighbours itself.  Only callers that know that exchange follows may
    // ask (a put without its exchange would leave words of the same exchange number behind).
    PutPlan make_put(int l, const Csr<TV>* /*unused*/ = nullptr) {
        if (l < 0 || l >= levels || !use_fused_put || !p2p.on || !comm.active() || !L[l].sp.dist || m != 1 ||
            !ctx.use_patterns || ctx.profiling)
            return no_put();
        return p2p.chan[l].put;
    }
    void note_put(int l, const PutPlan& pp, const TV* out) {
        if (!pp.on) return;
        MGB_CHECK(put_done[l] == nullptr, "fused put: the previous put of this channel was never exchanged");
        put_done[l] = out;
    }
    // y = op(M v) for a level-(lin+1) input vector v whose ghost rows must be refreshed first (row-partitioned
    // levels).  With the peer-memory exchange (p2p.cuh) the refresh runs on a second, high-priority stream BESIDE
    // the rows that read no ghost; the few rows next to the slab ends follow once both have finished.  The fork and
    // the join are event dependencies, so the overlap is part of the captured cycle graph.
    static constexpr int OVERLAP_HALO_CTAS = 64;   // x 256 threads: the room the interior kernel leaves free
    template <typename TA>
    void apply_x(int lin, const Csr<TA>& M, int mode, TV* v, const TV* b, const TV* d, TV* y, int kind, int level,
                 const TV* dpat = nullptr, int put_level = -1) {
        PutPlan pp = (put_level >= 0 && pattern_in_use(ctx, M, m)) ? make_put(put_level) : no_put();
        const bool need = comm.active() && lin >= 0 && lin < levels && L[lin].sp.dist;
        const bool split = need && p2p.on && ctx.use_overlap && !ctx.profiling && pattern_in_use(ctx, M, m) &&
                           2LL * (M.int_hi - M.int_lo) >= M.n_rows;
        if constexpr (std::is_same<TA, TV>::value) {
            // box-stencil kernel beside the exchange of its input vector: the kernel starts at once, its tiles that read
            // ghost rows come last and wait in the kernel for the exchange number (box.cuh, BoxWait)
            if (need && p2p.on && ctx.overlap_box && !ctx.profiling && m == 1 && M.box.ok && pattern_in_use(ctx, M, m) &&
                ctx.split_test == 0 && ctx.use_box && M.n_rows >= ctx.box_min_rows && mode != MODE_ADD && v != y) {
                MGB_CUDA(cudaEventRecord(ctx.ev_fork, ctx.stream));
                MGB_CUDA(cudaStreamWaitEvent(ctx.side, ctx.ev_fork, 0));
                exchange(lin, v, ctx.side, 0, true);
                MGB_CUDA(cudaEventRecord(ctx.ev_join, ctx.side));
                BoxWait bw;
                bw.epoch = p2p.epoch + lin;
                bw.consumed = p2p.consumed + lin;
                bw.ticket = p2p.ticket2 + lin;
                bool ran;
                {
                    const double fmt = M.pat.matrix_bytes(M.n_rows) + vec_bytes<TA, TV>(M, mode, m, dpat != nullptr);
                    Launch La(ctx, kind, level, csr_bytes<TA, TV>(M, mode, m), fmt);
                    ran = launch_box<TA, TV>(ctx, M, mode, v, b, d, dpat, y, pp, false, bw);
                    if (!ran) La.cancel();
                }
                MGB_CUDA(cudaStreamWaitEvent(ctx.stream, ctx.ev_join, 0));
                if (!ran) {      // the box kernel declined: mark the exchange consumed and run the ordinary pass
                    p2p_mark_consumed_kernel<<<1, 1, 0, ctx.stream>>>(p2p.epoch + lin, p2p.consumed + lin);
                    MGB_LAUNCH_CHECK();
                    csr_apply<TA, TV>(ctx, M, mode, v, b, d, y, m, kind, level, dpat, pp);
                }
                note_put(put_level, pp, y);
                return;
            }
        }
        if (!split) {
            if (need) exchange(lin, v);
            csr_apply<TA, TV>(ctx, M, mode, v, b, d, y, m, kind, level, dpat, pp);
            note_put(put_level, pp, y);
            return;
        }
        MGB_CUDA(cudaEventRecord(ctx.ev_fork, ctx.stream));
        MGB_CUDA(cudaStreamWaitEvent(ctx.side, ctx.ev_fork, 0));
        exchange(lin, v, ctx.side, OVERLAP_HALO_CTAS);
        MGB_CUDA(cudaEventRecord(ctx.ev_join, ctx.side));
        csr_apply_split<TA, TV>(ctx, M, mode, v, b, d, y, kind, level, dpat, M.int_lo, M.int_hi, OVERLAP_HALO_CTAS * 256,
                                [&] { MGB_CUDA(cudaStreamWaitEvent(ctx.stream, ctx.ev_join, 0)); }, pp);
        note_put(put_level, pp, y);
    }
    // reductions over a distributed level are completed by an in-place all-reduce on the stream
    void allreduce(int l, double* dptr, int k) {
        if (!comm.active() || !L[l].sp.dist) return;
        Launch La(ctx, K_REDUCE, l + 1, 0.0);
        MGB_NCCL(nccl().AllReduce(dptr, dptr, (size_t)k, ncclDouble, ncclSum, comm.comm, ctx.stream));
    }
    double norm(long long n, const TV* x, int l = 0) {
        dev_norm2sq<TV>(ctx, n, x, ctx.scal + 8);
        allreduce(l, ctx.scal + 8, 1);
        double v;
        read_scalars(ctx, ctx.scal + 8, 1, &v);
        return std::sqrt(v);
    }
    zc dot(long long n, const TV* x, const TV* y, int l = 0) {
        dev_dot<TV>(ctx, n, x, y, ctx.scal + 10);
        allreduce(l, ctx.scal + 10, 2);
        double v[2];
        read_scalars(ctx, ctx.scal + 10, 2, v);
        return zc(v[0], v[1]);
    }

    // halo exchange of a level-l vector laid out [ghosts below | owned | ghosts above], v at the first owned row
    void exchange(int l, TV* v, cudaStream_t stream = nullptr, int max_ctas = 0, bool consumer_waits = false) {
        if (!comm.active() || l < 0 || l >= levels || !L[l].sp.dist) return;
        DistSpace& sp = L[l].sp;
        if (!stream) stream = ctx.stream;
        MGB_CHECK(stream == ctx.stream || p2p.on, "only the peer-memory exchange runs on a side stream");
        Launch La(ctx, K_COPY, l + 1, 2.0 * sp.n_ghost * m * sizeof(TV));
        int skip_put = 0;
        if (put_done[l]) {
            MGB_CHECK(p2p.on && put_done[l] == v, "fused put: another vector is exchanged than the one that was put");
            put_done[l] = nullptr;
            skip_put = 1;
        }
        if (p2p.on) {
            // put into the neighbours' receive buffers over NVLink, then wait for theirs and unpack (p2p.cuh)
            ChanDev<TV>* cd = static_cast<ChanDev<TV>*>(p2p.chan[l].dev);
            const long long work = std::max<long long>((long long)sp.n_send, sp.n_ghost) * m;
            // no CTA of this kernel waits on another one (they poll words written by the peers), so the grid may
            // be as wide as the copy needs
            int g = (int)std::max<long long>(1, std::min<long long>((work + 511) / 512, 2LL * ctx.sm_count));
            if (max_ctas > 0) g = std::min(g, max_ctas);
            p2p_halo_kernel<TV><<<g, 256, 0, stream>>>(cd, v, sp.d_send_idx, sp.n_send, sp.n_ghost, sp.n_lo,
                                                            sp.n_owned, m, p2p.epoch + l, p2p.ticket + l,
                                                            p2p.trace ? p2p.trace + (size_t)l * P2P_TRACE_ROWS * 4 : nullptr,
                                                            skip_put, consumer_waits ? nullptr : p2p.consumed + l);
            MGB_LAUNCH_CHECK();
            return;
        }
        TV* sb = static_cast<TV*>(sp.sendbuf);
        if (sp.n_send > 0) {
            pack_kernel<TV><<<ctx.ew_blocks((long long)sp.n_send * m), 256, 0, ctx.stream>>>(v, sp.d_send_idx, sp.n_send, m, sb);
            MGB_LAUNCH_CHECK();
        }
        const size_t per = (size_t)m * (sizeof(TV) / sizeof(double));
        MGB_NCCL(nccl().GroupStart());
        for (int p = 0; p < comm.world; ++p) {
            if (sp.send_cnt[p] > 0)
                MGB_NCCL(nccl().Send(sb + (size_t)sp.send_off[p] * m, sp.send_cnt[p] * per, ncclDouble, p, comm.comm, ctx.stream));
            if (sp.recv_cnt[p] > 0)
                MGB_NCCL(nccl().Recv(v + sp.ghost_pos(sp.recv_off[p]) * m, sp.recv_cnt[p] * per, ncclDouble, p, comm.comm, ctx.stream));
        }
        MGB_NCCL(nccl().GroupEnd());
    }
    // assemble a replicated level-l vector from the owned pieces computed by each rank
    void allgather_rows(int l, TV* v, const std::vector<long long>& offs) {
        if (!comm.active()) return;
        Launch La(ctx, K_COPY, l + 1, 1.0 * L[l].n * m * sizeof(TV));
        if (p2p.on && p2p.gather_level == l) {
            ChanDev<TV>* cd = static_cast<ChanDev<TV>*>(p2p.chan[levels].dev);
            const long long off = offs[comm.rank] * m, cnt = (offs[comm.rank + 1] - offs[comm.rank]) * m;
            const int g = (int)std::max<long long>(1, std::min<long long>((L[l].n * m + 511) / 512, 2LL * ctx.sm_count));
            p2p_gather_kernel<TV><<<g, 256, 0, ctx.stream>>>(cd, v, off, cnt, m, p2p.epoch + levels, p2p.ticket + levels);
            MGB_LAUNCH_CHECK();
            return;
        }
        const size_t per = (size_t)m * (sizeof(TV) / sizeof(double));
        MGB_NCCL(nccl().GroupStart());
        for (int p = 0; p < comm.world; ++p) {
            const long long cnt = offs[p + 1] - offs[p];
            if (cnt > 0)
                MGB_NCCL(nccl().Broadcast(v + (size_t)offs[p] * m, v + (size_t)offs[p] * m, cnt * per, ncclDouble, p, comm.comm, ctx.stream));
        }
        MGB_NCCL(nccl().GroupEnd());
    }

    // ---- peer-memory exchange (p2p.cuh): IPC block, channel tables -------------------------------------
    void p2p_release() {
        if (!p2p.block && p2p.peer.empty()) return;
        cudaSetDevice(ctx.device);
        if (ctx.stream) cudaStreamSynchronize(ctx.stream);
        p2p_trace_report();
        for (int q = 0; q < (int)p2p.peer.size(); ++q)
            if (q != comm.rank && p2p.peer[q] && !(q < (int)p2p.same_process.size() && p2p.same_process[q]))
                cudaIpcCloseMemHandle(p2p.peer[q]);
        p2p.peer.clear();
        p2p.same_process.clear();
        for (auto& c : p2p.chan) {
            if (c.dev) cudaFree(c.dev);
            c.dev = nullptr;
        }
        p2p.chan.clear();
        // The exported block itself is NOT freed here: a peer may still have it mapped (ranks tear down at
        // different times and cudaFree of memory that is still imported elsewhere is undefined).  It is a few
        // MB and goes away with the process.
        p2p.block = nullptr;
        p2p.block_bytes = 0;
        dev_free(p2p.epoch);
        dev_free(p2p.ticket);
        dev_free(p2p.consumed);
        dev_free(p2p.ticket2);
        p2p.on = false;
        p2p.gather_level = -1;
    }
    // MGB200_P2P_TRACE=1: phase times of the last exchanges of every halo channel, printed at teardown
    void p2p_trace_report() {
        if (!p2p.trace) return;
        const int nchan = levels + 1;
        std::vector<unsigned long long> h((size_t)nchan * P2P_TRACE_ROWS * 4);
        cudaMemcpy(h.data(), p2p.trace, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
        for (int l = 0; l < levels; ++l) {
            double put = 0, wait = 0, unpack = 0;
            int cnt = 0;
            for (int k = 0; k < P2P_TRACE_ROWS; ++k) {
                const unsigned long long* t = h.data() + ((size_t)l * P2P_TRACE_ROWS + k) * 4;
                if (!t[0] || !t[3] || t[3] < t[0]) continue;
                put += (double)(t[1] - t[0]);
                wait += (double)(t[2] - t[1]);
                unpack += (double)(t[3] - t[2]);
                ++cnt;
            }
            if (cnt)
                std::fprintf(stderr, "[mgb200 p2p trace] rank %d level %d: %d exchanges, put %.2f us, poll + unpack %.2f us "
                             "(CTA 0)\n", comm.rank, l + 1, cnt, put / cnt * 1e-3, (wait + unpack) / cnt * 1e-3);
        }
        dev_free(p2p.trace);
    }
    void nccl_allgather_bytes(const void* mine, size_t bytes, std::vector<unsigned char>& all) {
        const int w = comm.world, r = comm.rank;
        unsigned char* d = dev_alloc<unsigned char>(bytes * w);
        MGB_CUDA(cudaMemcpyAsync(d + bytes * r, mine, bytes, cudaMemcpyHostToDevice, ctx.stream));
        MGB_NCCL(nccl().AllGather(d + bytes * r, d, bytes, ncclInt8, comm.comm, ctx.stream));
        all.resize(bytes * w);
        MGB_CUDA(cudaMemcpyAsync(all.data(), d, bytes * w, cudaMemcpyDeviceToHost, ctx.stream));
        ctx.sync();
        dev_free(d);
    }
    // collective over all ranks (called from ensure_work)
    void p2p_setup() {
        p2p_release();
        if (!comm.active() || env_int("MGB200_P2P", 1) == 0 || comm.world > P2P_MAXW) return;
        const int w = comm.world, r = comm.rank;
        const int nchan = levels + 1;
        auto align = [](size_t v) { return (v + 255) / 256 * 256; };
        p2p.chan.assign(nchan, ChanHost());
        size_t off = align((size_t)nchan * P2P_MAXW * sizeof(unsigned long long));
        for (int l = 0; l < levels - 1; ++l) {
            if (!L[l].sp.dist) continue;
            p2p.chan[l].used = true;
            p2p.chan[l].rows = L[l].sp.n_ghost;
            if (!L[l + 1].sp.dist) {
                p2p.gather_level = l + 1;
                p2p.chan[levels].used = true;
                p2p.chan[levels].rows = L[l].nc_global;
            }
        }
        for (auto& c : p2p.chan) {
            if (!c.used) continue;
            for (int par = 0; par < 2; ++par) {
                c.buf_off[par] = off;
                off += align((size_t)std::max<long long>(c.rows, 1) * m * sizeof(TV) * 2);   // LL words: 2x the data
            }
        }
        p2p.block_bytes = off;
        int ok = 1;
        cudaIpcMemHandle_t mine;
        std::memset(&mine, 0, sizeof(mine));
        if (cudaMalloc(&p2p.block, p2p.block_bytes) != cudaSuccess) {
            p2p.block = nullptr;
            ok = 0;
        } else {
            MGB_CUDA(cudaMemset(p2p.block, 0, p2p.block_bytes));
            if (cudaIpcGetMemHandle(&mine, p2p.block) != cudaSuccess) ok = 0;
        }
        cudaGetLastError();
        // table: handle | ok | process id | device | block address | per channel {buf_off[2], recv_off[w]}
        const size_t per_chan = 2 + (size_t)w;
        std::vector<long long> tab(8 + 4 + per_chan * nchan, 0);
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
        std::memcpy(tab.data(), &mine, 64);
        tab[8] = ok;
        tab[9] = (long long)getpid();
        tab[10] = ctx.device;
        tab[11] = (long long)reinterpret_cast<uintptr_t>(p2p.block);
        for (int c = 0; c < nchan; ++c) {
            long long* t = tab.data() + 12 + per_chan * c;
            t[0] = (long long)p2p.chan[c].buf_off[0];
            t[1] = (long long)p2p.chan[c].buf_off[1];
            if (c < levels && p2p.chan[c].used)
                for (int q = 0; q < w; ++q) t[2 + q] = L[c].sp.recv_off[q];
        }
        std::vector<unsigned char> allb;
        nccl_allgather_bytes(tab.data(), tab.size() * sizeof(long long), allb);
        const long long* all = reinterpret_cast<const long long*>(allb.data());
        auto T = [&](int q) { return all + (size_t)q * tab.size(); };
        for (int q = 0; q < w; ++q) ok = ok && (int)T(q)[8];
        p2p.peer.assign(w, nullptr);
        p2p.same_process.assign(w, 0);
        if (ok) {
            p2p.peer[r] = p2p.block;
            for (int q = 0; q < w && ok; ++q) {
                if (q == r) continue;
                if (T(q)[9] == (long long)getpid()) {
                    // a device driven by another thread of this process (mgb200_multi_*): plain peer access
                    int can = 0;
                    const int dq = (int)T(q)[10];
                    if (cudaDeviceCanAccessPeer(&can, ctx.device, dq) != cudaSuccess || !can) {
                        ok = 0;
                        cudaGetLastError();
                    } else {
                        const cudaError_t e = cudaDeviceEnablePeerAccess(dq, 0);
                        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ok = 0;
                        cudaGetLastError();
                        p2p.peer[q] = reinterpret_cast<unsigned char*>((uintptr_t)T(q)[11]);
                        p2p.same_process[q] = 1;
                    }
                    continue;
                }
                cudaIpcMemHandle_t hq;
                std::memcpy(&hq, T(q), 64);
                void* ptr = nullptr;
                if (cudaIpcOpenMemHandle(&ptr, hq, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                    ok = 0;
                    cudaGetLastError();
                } else {
                    p2p.peer[q] = static_cast<unsigned char*>(ptr);
                }
            }
        }
        // every rank must reach the same verdict (and nobody may publish before all blocks are zeroed)
        {
            int* dflag = dev_alloc<int>(1);
            MGB_CUDA(cudaMemcpyAsync(dflag, &ok, sizeof(int), cudaMemcpyHostToDevice, ctx.stream));
            MGB_NCCL(nccl().AllReduce(dflag, dflag, 1, ncclInt32, ncclMin, comm.comm, ctx.stream));
            MGB_CUDA(cudaMemcpyAsync(&ok, dflag, sizeof(int), cudaMemcpyDeviceToHost, ctx.stream));
            ctx.sync();
            dev_free(dflag);
        }
        if (!ok) {
            p2p_release();
            return;
        }
        for (int c = 0; c < nchan; ++c) {
            ChanHost& ch = p2p.chan[c];
            if (!ch.used) continue;
            ChanDev<TV> cd;
            std::memset(static_cast<void*>(&cd), 0, sizeof(cd));
            cd.world = w;
            cd.rank = r;
            const bool gather = (c == levels);
            const Level<TV>& lg = gather ? L[p2p.gather_level - 1] : L[c];
            for (int par = 0; par < 2; ++par) cd.rbuf[par] = reinterpret_cast<const TV*>(p2p.block + ch.buf_off[par]);
            cd.send_off[0] = 0;
            for (int q = 0; q < w; ++q) {
                const long long* tq = T(q) + 12 + per_chan * c;
                const long long land = gather ? lg.coarse_row_offsets[r] : tq[2 + r];
                for (int par = 0; par < 2; ++par)   // landing zone in LL words (2 x sizeof(TV) bytes per element)
                    cd.dst[par][q] = reinterpret_cast<TV*>(p2p.peer[q] + tq[par] + (size_t)land * m * sizeof(TV) * 2);
                cd.flag_dst[q] = reinterpret_cast<unsigned long long*>(p2p.peer[q]) + (size_t)c * P2P_MAXW + r;
                cd.flag_src[q] = reinterpret_cast<const unsigned long long*>(p2p.block) + (size_t)c * P2P_MAXW + q;
                if (gather) {
                    cd.send_off[q + 1] = 0;
                    cd.recv_off[q] = (int)lg.coarse_row_offsets[q];
                    cd.recv_cnt[q] = (int)(lg.coarse_row_offsets[q + 1] - lg.coarse_row_offsets[q]);
                } else {
                    cd.send_off[q + 1] = cd.send_off[q] + lg.sp.send_cnt[q];
                    cd.recv_off[q] = lg.sp.recv_off[q];
                    cd.recv_cnt[q] = lg.sp.recv_cnt[q];
                }
            }
            MGB_CUDA(cudaMalloc(&ch.dev, sizeof(cd)));
            MGB_CUDA(cudaMemcpy(ch.dev, &cd, sizeof(cd), cudaMemcpyHostToDevice));
            // fused put (ll.cuh): possible when the rows the neighbours asked for are exactly the two end ranges of
            // the owned rows, in order (z-slabs: first plane to the rank below, last plane to the rank above)
            ch.put = no_put();
            if (!gather && m == 1) {
                const DistSpace& sp = lg.sp;
                bool ok = true;
                for (int q = 0; q < w; ++q)
                    if (sp.send_cnt[q] > 0 && q != r - 1 && q != r + 1) ok = false;
                const int lo_cnt = r > 0 ? sp.send_cnt[r - 1] : 0, hi_cnt = r + 1 < w ? sp.send_cnt[r + 1] : 0;
                const long long hi_start = sp.n_owned - hi_cnt;
                if (ok && (int)sp.send_idx_host.size() == sp.n_send) {
                    for (int k = 0; k < lo_cnt && ok; ++k) ok = sp.send_idx_host[sp.send_off[r - 1] + k] == k;
                    for (int k = 0; k < hi_cnt && ok; ++k) ok = sp.send_idx_host[sp.send_off[r + 1] + k] == hi_start + k;
                } else {
                    ok = false;
                }
                if (ok && lo_cnt + hi_cnt > 0) {
                    ch.put.on = 1;
                    ch.put.lo_cnt = lo_cnt;
                    ch.put.hi_cnt = hi_cnt;
                    ch.put.hi_start = hi_cnt > 0 ? (int)hi_start : 0x7fffffff;
                    for (int par = 0; par < 2; ++par) {
                        ch.put.dst_lo[par] = lo_cnt > 0 ? reinterpret_cast<unsigned long long*>(cd.dst[par][r - 1]) : nullptr;
                        ch.put.dst_hi[par] = hi_cnt > 0 ? reinterpret_cast<unsigned long long*>(cd.dst[par][r + 1]) : nullptr;
                    }
                }
            }
        }
        if (env_int("MGB200_P2P_TRACE", 0)) {
            p2p.trace = dev_alloc<unsigned long long>((size_t)nchan * P2P_TRACE_ROWS * 4);
            MGB_CUDA(cudaMemset(p2p.trace, 0, (size_t)nchan * P2P_TRACE_ROWS * 4 * sizeof(unsigned long long)));
        }
        p2p.epoch = dev_alloc<unsigned long long>(nchan);
        p2p.ticket = dev_alloc<unsigned>(nchan);
        p2p.consumed = dev_alloc<unsigned long long>(nchan);
        p2p.ticket2 = dev_alloc<unsigned>(nchan);
        MGB_CUDA(cudaMemset(p2p.epoch, 0, nchan * sizeof(unsigned long long)));
        MGB_CUDA(cudaMemset(p2p.ticket, 0, nchan * sizeof(unsigned)));
        MGB_CUDA(cudaMemset(p2p.consumed, 0, nchan * sizeof(unsigned long long)));
        MGB_CUDA(cudaMemset(p2p.ticket2, 0, nchan * sizeof(unsigned)));
        for (int c = 0; c < levels; ++c) p2p.chan[c].put.epoch = p2p.epoch + c;
        put_done.assign(levels, nullptr);
        p2p.on = true;
    }
    static TV to_tv(zc a) { return VT<TV>::make(a.real(), a.imag()); }

    // relax (MGcycle.jl:122-136) fused: each sweep is  x' = x + d.*(b - A x).  With x known to be
    // zero the first sweep is x = d.*b.  Returns the buffer holding the result (x or scratch).
    // last_exchanged: the caller knows that the vector returned is the next one exchanged on channel l
    // (fused put, ll.cuh); the intermediate iterates always are (by the next sweep).
    TV* relax(int l, const TV* b, TV* x, TV* scratch, int numit, bool xzero, bool last_exchanged = false) {
        Level<TV>& lv = L[l];
        int sweeps = std::max(numit, 1);  // numit = 0 still does one update (MGcycle.jl:134)
        const TV* dpat = (lv.dpat && m == 1 && ctx.use_patterns) ? lv.dpat : nullptr;
        if (xzero && dpat && sweeps >= 2 && ctx.fuse_first_sweeps && !lv.sp.dist) {
            // x1 = d .* b and x2 = x1 + d .* (b - A x1) in one pass of the box-stencil kernel (box.cuh, MODE 4)
            const double bytes = 3.0 * lv.n * sizeof(TV) + csr_bytes<TV, TV>(lv.A, MODE_SWEEP, 1);
            const double fmt = lv.n * (2.0 * sizeof(TV) + 2.0);
            Launch La(ctx, K_FIRST2, l + 1, bytes, fmt);
            if (launch_box<TV, TV>(ctx, lv.A, MODE_SWEEP2_FROM_ZERO, b, nullptr, nullptr, dpat, scratch, no_put())) {
                std::swap(x, scratch);
                sweeps -= 2;
                xzero = false;
            } else {
                La.cancel();
            }
        }
        if (xzero) {
            if (dpat) {
                const PutPlan pp = (sweeps > 1 || last_exchanged) ? make_put(l) : no_put();
                Launch La(ctx, K_DIAG, l + 1, 3.0 * lv.n * sizeof(TV), lv.n * (2.0 * sizeof(TV) + 2.0));
                diag_scale_pat_kernel<TV><<<ctx.ew_blocks(lv.n), 256, 0, ctx.stream>>>(pp, lv.n, lv.A.pat.pid, dpat, b, x);
                note_put(l, pp, x);
            } else {
                Launch La(ctx, K_DIAG, l + 1, (2.0 * m + 1.0) * lv.n * sizeof(TV));
                diag_scale_kernel<TV><<<ctx.ew_blocks(lv.n * m), 256, 0, ctx.stream>>>(lv.n, m, lv.d, b, x);
            }
            MGB_LAUNCH_CHECK();
            sweeps -= 1;
        }
        for (int s = 0; s < sweeps; ++s) {
            const bool put = (s + 1 < sweeps) || last_exchanged;
            apply_x<TV>(l, lv.A, MODE_SWEEP, x, b, lv.d, scratch, K_SWEEP, l + 1, dpat, put ? l : -1);
            std::swap(x, scratch);
        }
        return x;
    }

    void solve_coarsest(const TV* b, TV* x) {  // x = A_L^{-1} b  (MGcycle.jl:176-179)
        const int n = coarse.n;
        if (coarse.kind == 1) {
            // coarseSolveType "GMRES" (MGcycle.jl:152-168): x .= 0; one restart of fgmres(10), tol 0.01,
            // right preconditioner M2(v) = d .* v, not flexible
            MGB_CHECK(m == 1, "coarsest GMRES: blockFGMRES (nrhs > 1) is not provided");
            dev_zero<TV>(ctx, n, x);
            FgmresWs ws;
            ws.r = coarse.r; ws.w = coarse.w; ws.t = coarse.t; ws.V = coarse.V; ws.Z = nullptr;
            ws.ld = n;
            PrecFn m2 = [this, n](const TV* v) -> TV* {
                Launch La(ctx, K_DIAG, levels, 3.0 * n * sizeof(TV));
                diag_scale_kernel<TV><<<ctx.ew_blocks(n), 256, 0, ctx.stream>>>(n, 1, coarse.d, v, coarse.dv);
                MGB_LAUNCH_CHECK();
                return coarse.dv;
            };
            int flag = 0, nres = 0;
            fgmres_core(L[levels - 1].A, levels, n, b, x, 10, false, 0.01, 1, m2, ws, &flag, nullptr, &nres);
            return;
        }
        typedef typename Wide<TV>::type TW;
        Launch La(ctx, K_COARSE, levels, (double)n * n * sizeof(TW));
        if (m == 1) {
            const int grid = cdiv((long long)n * 32, 256);
            lower_apply_kernel<TW, TV, 1><<<grid, 256, 0, ctx.stream>>>(n, m, coarse.linv, coarse.perm, b, coarse.y);
            upper_apply_kernel<TW, TV, 1><<<grid, 256, 0, ctx.stream>>>(n, m, coarse.uinv, coarse.y, x);
        } else {
            const int grid = cdiv((long long)n * cdiv(m, AP_MC) * 32, 256);
            lower_apply_kernel<TW, TV, AP_MC><<<grid, 256, 0, ctx.stream>>>(n, m, coarse.linv, coarse.perm, b, coarse.y);
            upper_apply_kernel<TW, TV, AP_MC><<<grid, 256, 0, ctx.stream>>>(n, m, coarse.uinv, coarse.y, x);
        }
        MGB_LAUNCH_CHECK();
    }

    typedef std::function<TV*(const TV*)> PrecFn;

    // FGMRES_relaxation (FGMRES.jl:48-126).  x0 += Z t.  Returns the number of inner steps done.
    int fgmres_relaxation(const Csr<TV>& A, int level, const TV* r0, TV* x0, int inner, const PrecFn& prec,
                          double TOL, FgmresMem<TV>& mem, long long n) {
        const long long nm = n * m;
        MGB_CHECK(mem.inner == inner, "FGMRES_relaxation: size of Krylov subspace is different than inner");
        MGB_CHECK(inner <= MAXK, "FGMRES_relaxation: inner too large");
        dev_zero<TV>(ctx, nm * inner, mem.Z);   // resetMem (FGMRES.jl:10-15)
        dev_zero<TV>(ctx, nm * inner, mem.AZ);
        const double rnorm0 = norm(nm, r0, level - 1);
        std::vector<zc> H((size_t)inner * inner, zc(0, 0)), xi(inner, zc(0, 0)), t(inner, zc(0, 0));
        const TV* w = nullptr;
        int done = 0;
        for (int j = 0; j < inner; ++j) {
            const TV* z = prec(j == 0 ? r0 : w);
            dev_copy<TV>(ctx, nm, z, mem.Z + (size_t)j * nm);
            apply_A(A, z, mem.Az, level);
            w = mem.Az;
            dev_copy<TV>(ctx, nm, w, mem.AZ + (size_t)j * nm);
            // t = AZ^H w over all `inner` columns (unused ones are zero), xi[j] = <w, r0>
            dev_multi_dot<TV>(ctx, nm, mem.AZ, nm, inner, w, ctx.scal + 16);
            dev_dot<TV>(ctx, nm, w, r0, ctx.scal + 16 + 2 * inner);
            allreduce(level - 1, ctx.scal + 16, 2 * inner + 2);
            std::vector<double> hv(2 * inner + 2);
            read_scalars(ctx, ctx.scal + 16, 2 * inner + 2, hv.data());
            for (int i = 0; i < inner; ++i) t[i] = zc(hv[2 * i], hv[2 * i + 1]);
            xi[j] = zc(hv[2 * inner], hv[2 * inner + 1]);
            for (int i = 0; i < inner; ++i) H[(size_t)i * inner + j] = t[i];
            for (int i = 0; i < inner; ++i) H[(size_t)j * inner + i] = std::conj(t[i]);
            std::vector<zc> Hs((size_t)inner * inner);
            for (int a = 0; a < inner; ++a)
                for (int c = 0; c < inner; ++c)
                    Hs[(size_t)a * inner + c] = 0.5 * H[(size_t)a * inner + c] + 0.5 * std::conj(H[(size_t)c * inner + a]);
            H = Hs;
            hermitian_pinv_apply(inner, H, xi, t, sizeof(typename VT<TV>::real_t) == 4 ? (double)std::numeric_limits<float>::epsilon()
                                                                                       : std::numeric_limits<double>::epsilon());
            zc tHt(0, 0), txi(0, 0);
            for (int a = 0; a < inner; ++a) {
                zc Ht(0, 0);
                for (int c = 0; c < inner; ++c) Ht += H[(size_t)a * inner + c] * t[c];
                tHt += std::conj(t[a]) * Ht;
                txi += std::conj(t[a]) * xi[a];
            }
            const double rn = std::sqrt(std::fabs((tHt - 2.0 * txi + rnorm0 * rnorm0).real()));
            done = j + 1;
            if (rn < TOL) break;
        }
        std::vector<TV> coef(inner);
        for (int i = 0; i < inner; ++i) coef[i] = to_tv(t[i]);
        dev_multi_axpy<TV>(ctx, nm, mem.Z, nm, inner, coef.data(), 1.0, x0, nullptr);  // x0 += Z t
        return done;
    }

    // recursiveCycle (MGcycle.jl:1-118).  l is 0-based.  x holds the iterate, scratch is its
    // ping-pong partner; returns the buffer with the result.  ctype is passed down explicitly
    // (the reference flips param.cycleType for the V leg of an F cycle, :82-84).
    TV* cycle(int l, const TV* b, TV* x, TV* scratch, bool xzero, char ctype) {
        if (l == levels - 1) {  // only when levels == 1 (MGcycle.jl:13-18)
            solve_coarsest(b, x);
            return x;
        }
        Level<TV>& lv = L[l];
        Level<TV>& lc = L[l + 1];
        const int npre = pre[l], npost = post[l];
        TV* xc_cur;
        // ---- pre-relaxation ----
        if (relax_kind == 1) {
            if (xzero) dev_zero<TV>(ctx, lv.n * m, x);
            if (xzero) dev_copy<TV>(ctx, lv.n * m, b, lv.r); else residual(lv.A, b, x, lv.r, l + 1);
            jac_gmres(l, lv.r, x, std::max(npre, 1));
        } else {
            TV* xn = relax(l, b, x, scratch, npre, xzero, true);   // the residual below exchanges the result
            if (xn != x) std::swap(x, scratch);
        }
        residual(lv.A, b, x, lv.r, l + 1, relax_kind == 0 ? l : -1);                // :58-60; r is exchanged for R
        if (lv.sp.dist && !lc.sp.dist) {
            // last distributed level: each rank restricts its own coarse rows, then the pieces are gathered
            apply_x<RT>(l, lv.R, MODE_SPMV, lv.r, nullptr, nullptr,
                            lc.b + (size_t)lv.coarse_row_offsets[comm.rank] * m, K_RESTRICT, l + 1);
            allgather_rows(l + 1, lc.b, lv.coarse_row_offsets);
        } else {
            apply_x<RT>(l, lv.R, MODE_SPMV, lv.r, nullptr, nullptr, lc.b, K_RESTRICT, l + 1);  // :66
        }
        if (l + 1 == levels - 1) {
            solve_coarsest(lc.b, lc.x0);                                           // :67-69
            xc_cur = lc.x0;
        } else if (ctype == 'K') {                                                 // :72-76
            FgmresMem<TV>& mk = lc.memK;
            dev_zero<TV>(ctx, lc.n * m, lc.x0);
            PrecFn mmg = [&, this](const TV* v) -> TV* { return cycle(l + 1, v, mk.vp0, mk.vp1, true, 'K'); };
            fgmres_relaxation(lc.A, l + 2, lc.b, lc.x0, 2, mmg, 1e-5, mk, lc.n);
            xc_cur = lc.x0;
        } else {
            xc_cur = cycle(l + 1, lc.b, lc.x0, lc.x1, true, ctype);                // :78
            if (ctype == 'W') {
                TV* other = (xc_cur == lc.x0) ? lc.x1 : lc.x0;
                xc_cur = cycle(l + 1, lc.b, xc_cur, other, false, 'W');            // :79-80
            } else if (ctype == 'F') {
                TV* other = (xc_cur == lc.x0) ? lc.x1 : lc.x0;
                xc_cur = cycle(l + 1, lc.b, xc_cur, other, false, 'V');            // :81-85
            }
        }
        // the corrected x is exchanged by the first post-sweep
        apply_x<RT>(l + 1, lv.P, MODE_ADD, xc_cur, nullptr, nullptr, x, K_PROLONG, l + 1, nullptr,
                        relax_kind == 0 ? l : -1);                                  // :90
        // ---- post-relaxation (:92-103) ----
        if (relax_kind == 1) {
            residual(lv.A, b, x, lv.r, l + 1);
            jac_gmres(l, lv.r, x, std::max(npost, 1));
            return x;
        }
        // below the finest level the result goes back to a parent that exchanges it next (its prolongation, or the
        // first sweep of the second visit of a W / F cycle); the K-cycle parent copies it around first
        return relax(l, b, x, scratch, npost, false, l >= 1 && ctype != 'K');
    }

    // "Jac-GMRES" smoother: FGMRES_relaxation with MM(v) = D.*v (MGcycle.jl:33-36,48-51,96-99)
    void jac_gmres(int l, const TV* r, TV* x, int inner) {
        Level<TV>& lv = L[l];
        FgmresMem<TV>& mem = lv.memRelax;
        // adjustMemoryForNumRHS sizes memRelax[l] for max(relaxPre(l), relaxPost(l)) (MGsetup.jl:209-211) and
        // FGMRES_relaxation raises an error when `inner` differs from that size (FGMRES.jl:60-62): same here
        const int maxRelax = std::max(std::max(pre[l], post[l]), 1);
        if (mem.inner != maxRelax) alloc_fgmres(mem, (size_t)lv.nalloc * m, maxRelax, vec_pad(l));
        PrecFn mm = [&, this](const TV* v) -> TV* {
            // y = D .* v  (0 + d*v is exact)
            Launch La(ctx, K_DIAG, l + 1, (2.0 * m + 1.0) * lv.n * sizeof(TV));
            diag_scale_kernel<TV><<<ctx.ew_blocks(lv.n * m), 256, 0, ctx.stream>>>(lv.n, m, lv.d, v, mem.vp0);
            MGB_LAUNCH_CHECK();
            return mem.vp0;
        };
        fgmres_relaxation(lv.A, l + 1, r, x, inner, mm, 1e-5, mem, lv.n);
    }

    // every fused put must have met its exchange by the end of a cycle (a put left behind would leave words that
    // carry the number of the NEXT exchange in the neighbours' buffers)
    void check_no_pending_put() {
        for (int l = 0; l < levels; ++l)
            MGB_CHECK(put_done[l] == nullptr, "fused put without its exchange at the end of a cycle");
    }
    // cycle from the finest level, replayed from a CUDA graph when the cycle has no host read-backs
    TV* cycle_fine(const TV* b, TV* x, TV* scratch, bool xzero, char ctype) {
        const bool can = ctx.use_graphs && !ctx.profiling && relax_kind == 0 && ctype != 'K' && coarse.kind == 0 &&
                         (!comm.active() || p2p.on) && levels > 1;
        if (!can) {
            TV* res = cycle(0, b, x, scratch, xzero, ctype);
            check_no_pending_put();
            return res;
        }
        const auto key = std::make_tuple(b, x, scratch, xzero, ctype);
        auto it = graphs.find(key);
        if (it == graphs.end()) {
            for (int l = 0; l < levels - 1; ++l) box_prepare<TV>(ctx, L[l].A);    // allocations stay outside the capture
            const long long l0 = ctx.launches;
            MGB_CUDA(cudaStreamBeginCapture(ctx.stream, cudaStreamCaptureModeThreadLocal));
            TV* res = nullptr;
            cudaGraph_t g = nullptr;
            try {
                res = cycle(0, b, x, scratch, xzero, ctype);
                check_no_pending_put();
            } catch (...) {
                cudaStreamEndCapture(ctx.stream, &g);
                if (g) cudaGraphDestroy(g);
                throw;
            }
            MGB_CUDA(cudaStreamEndCapture(ctx.stream, &g));
            GraphEntry e;
            e.result = res;
            e.launches = ctx.launches - l0;
            ctx.launches = l0;
            cudaError_t err = cudaGraphInstantiate(&e.exec, g, 0);
            cudaGraphDestroy(g);
            MGB_CUDA(err);
            it = graphs.emplace(key, e).first;
        }
        MGB_CUDA(cudaGraphLaunch(it->second.exec, ctx.stream));
        ctx.launches += it->second.launches;
        return it->second.result;
    }

    // one cycle on the caller's buffers (b in L[0].b, x in ucur); the result pointer is returned
    TV* cycle_top(bool xzero) {
        MGB_CHECK(!krylov_only, "a mixed-precision handle has no cycle of its own: call the single-precision handle");
        ensure_work();
        TV* other = (ucur == ux0) ? ux1 : ux0;
        ucur = cycle_fine(L[0].b, ucur, other, xzero, cycle_type);
        return ucur;
    }

    // ---- solveMG (SolveFuncs.jl:3-39) on the device buffers L[0].b / xcur ---------------------
    int solveMG(double tol, int max_iter, double* resvec) {
        MGB_CHECK(!krylov_only, "solveMG on a mixed-precision handle: call the single-precision handle");
        ensure_work();
        Level<TV>& lv = L[0];
        const long long nm = lv.n * m;
        TV*& xcur = ucur;
        bool xzero = (norm(nm, xcur) == 0.0);
        double res;
        if (xzero) {
            res = norm(nm, lv.b);
        } else {
            residual(lv.A, lv.b, xcur, lv.r, 1);
            res = norm(nm, lv.r);
        }
        const double res_init = res;
        resvec[0] = res_init;
        int iter = 0;
        for (int count = 1; count <= max_iter; ++count) {
            TV* other = (xcur == ux0) ? ux1 : ux0;
            xcur = cycle_fine(lv.b, xcur, other, xzero, cycle_type);
            xzero = false;
            residual(lv.A, lv.b, xcur, lv.r, 1);
            iter += 1;
            res = norm(nm, lv.r);
            resvec[count] = res;
            if (res / res_init < tol) break;
        }
        return iter;
    }

    // getMultigridPreconditioner (SolveFuncs.jl:43-63): z .= 0; recursiveCycle(param,r,z,1); z
    TV* precondition(const TV* r) {
        if (ext_prec) return ext_prec(r);
        Level<TV>& lv = L[0];
        return cycle_fine(r, lv.x0, lv.x1, true, cycle_type);
    }

    // ---- KrylovMethods.cg with M = one cycle (solveCG_MG, SolveFuncs.jl:103-116) --------------
    // b: L[0].b, x: in/out buffer xk (n).  resvec[max_iter].  Returns iterations; flag out.
    int solveCG(TV* xk, double tol, int max_iter, int* flag, double* resvec) {
        ensure_work();
        MGB_CHECK(m == 1, "solveCG: block CG (nrhs > 1) goes through solveBlockCG");
        MGB_CHECK(!VT<TV>::is_complex, "KrylovMethods.cg compares alpha < 0 and is real-only");
        ensure_krylov(0, false);
        Level<TV>& lv = L[0];
        const long long n = lv.n;
        const Csr<TV>& A = krylov_A();
        const TV* b = lv.b;
        const double nb = norm(n, b);
        if (nb == 0.0) {
            dev_zero<TV>(ctx, n, xk);
            *flag = -9;
            resvec[0] = 0.0;
            return 0;
        }
        residual(A, b, xk, kr, 1);                       // r = b - A(x)
        TV* z = precondition(kr);                        // z = M(r)
        dev_copy<TV>(ctx, n, z, kp);                     // p = copy(z)
        double* s = ctx.scal;                            // s[0]=gamma s[2]=delta s[4]=rr s[6]=gamma_new
        dev_dot<TV>(ctx, n, kr, z, s + 0);
        allreduce(0, s + 0, 2);
        *flag = -1;
        int last = 0;
        for (int it = 1; it <= max_iter; ++it) {
            last = it;
            apply_A(A, kp, kAp, 1);                      // Ap = A(p)
            dev_dot<TV>(ctx, n, kp, kAp, s + 2);         // delta = <p,Ap>   (gamma = <r,z> already in s[0])
            allreduce(0, s + 2, 2);
            {
                Launch La(ctx, K_VECTOR, 0, 6.0 * n * sizeof(TV));
                cg_update_kernel<TV><<<ctx.red_blocks(n), RED_THREADS, 0, ctx.stream>>>(n, s, kp, kAp, xk, kr, ctx.red, s + 4);
                MGB_LAUNCH_CHECK();
            }
            allreduce(0, s + 4, 1);
            double hv[6];
            read_scalars(ctx, s, 6, hv);
            const double alpha = hv[0] / hv[2];
            if (std::isinf(alpha) || alpha < 0.0) {
                *flag = -2;
                break;
            }
            resvec[it - 1] = std::sqrt(hv[4]) / nb;
            if (resvec[it - 1] <= tol) {
                *flag = 0;
                break;
            }
            z = precondition(kr);
            dev_dot<TV>(ctx, n, z, kr, s + 6);           // <z,r>
            allreduce(0, s + 6, 2);
            {
                Launch La(ctx, K_VECTOR, 0, 3.0 * n * sizeof(TV));
                cg_direction_kernel<TV><<<ctx.ew_blocks(n), 256, 0, ctx.stream>>>(n, s + 6, s + 0, z, kp);
                MGB_LAUNCH_CHECK();
            }
            // gamma for the next iteration: dot(r,z) == dot(z,r) bit for bit in real arithmetic
            MGB_CUDA(cudaMemcpyAsync(s + 0, s + 6, 2 * sizeof(double), cudaMemcpyDeviceToDevice, ctx.stream));
        }
        return last;
    }

    // ---- KrylovMethods.blockCG with M = one cycle (solveCG_MG with nrhs > 1, SolveFuncs.jl:113) ----
    // O'Leary block CG on n x m blocks (RHS-fastest on the device).  resmat: max_iter x m (row-major).
    double* gram_ws = nullptr;      // device: 3 Gram results (2 m^2 doubles each) + m x m coefficient matrix
    double* gram_partials = nullptr;
    unsigned* gram_counter = nullptr;
    TV* kq = nullptr;               // spare n x m block (P ping-pong)
    void gram(long long n, const TV* X, const TV* Y, double* out) {
        Launch La(ctx, K_REDUCE, 0, 2.0 * n * m * sizeof(TV));
        const int blocks = (int)std::max<long long>(1, std::min<long long>((n + GRAM_ROWS - 1) / GRAM_ROWS,
                                                                           std::min(GRAM_MAX_BLOCKS, ctx.sm_count * 4)));
        const size_t smem = 2 * (size_t)GRAM_ROWS * m * sizeof(TV);
        if (smem > 48 * 1024)
            MGB_CUDA(cudaFuncSetAttribute(gram_kernel<TV>, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx.max_smem_optin));
        gram_kernel<TV><<<blocks, GRAM_THREADS, smem, ctx.stream>>>(n, m, X, Y, gram_partials, gram_counter, out);
        MGB_LAUNCH_CHECK();
        allreduce(0, out, 2 * m * m);
    }
    void read_matrix(const double* dptr, std::vector<zc>& M) {
        std::vector<double> h(2 * (size_t)m * m);
        MGB_CUDA(cudaMemcpyAsync(h.data(), dptr, h.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
        ctx.sync();
        M.resize((size_t)m * m);
        for (int i = 0; i < m * m; ++i) M[i] = zc(h[2 * i], h[2 * i + 1]);
    }
    // out = base + X * C   (C: m x m host matrix)
    void block_axpy(long long n, const TV* X, const std::vector<zc>& C, const TV* base, TV* out) {
        std::vector<TV> hc((size_t)m * m);
        for (int i = 0; i < m * m; ++i) hc[i] = to_tv(C[i]);
        TV* dC = reinterpret_cast<TV*>(gram_ws + 6 * (size_t)m * m);
        MGB_CUDA(cudaMemcpyAsync(dC, hc.data(), hc.size() * sizeof(TV), cudaMemcpyHostToDevice, ctx.stream));
        ctx.sync();  // hc is a stack temporary
        Launch La(ctx, K_VECTOR, 0, 3.0 * n * m * sizeof(TV));
        if ((size_t)m * m * sizeof(TV) > 48 * 1024)
            MGB_CUDA(cudaFuncSetAttribute(block_axpy_kernel<TV>, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx.max_smem_optin));
        block_axpy_kernel<TV><<<ctx.ew_blocks(n * m), 256, (size_t)m * m * sizeof(TV), ctx.stream>>>(n, m, X, dC, base, out);
        MGB_LAUNCH_CHECK();
    }
    void colnorms(long long n, const TV* R, std::vector<double>& out) {
        {
            Launch La(ctx, K_REDUCE, 0, 1.0 * n * m * sizeof(TV));
            colnorm2_kernel<TV><<<ctx.red_blocks(n * m), RED_THREADS, 0, ctx.stream>>>(n, m, R, gram_partials, gram_counter, gram_ws);
            MGB_LAUNCH_CHECK();
        }
        allreduce(0, gram_ws, m);
        out.resize(m);
        MGB_CUDA(cudaMemcpyAsync(out.data(), gram_ws, m * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
        ctx.sync();
        for (auto& v : out) v = std::sqrt(v);
    }
    static void matmul_small(int m, const std::vector<zc>& A, const std::vector<zc>& B, std::vector<zc>& C, double sgn) {
        C.assign((size_t)m * m, zc(0, 0));
        for (int i = 0; i < m; ++i)
            for (int k = 0; k < m; ++k) {
                const zc a = A[(size_t)i * m + k];
                for (int j = 0; j < m; ++j) C[(size_t)i * m + j] += sgn * a * B[(size_t)k * m + j];
            }
    }
    void ensure_block_ws() {
        if (!gram_ws) {
            gram_ws = dev_alloc<double>(6 * (size_t)64 * 64 + 2 * (size_t)64 * 64 + 64);
            gram_partials = dev_alloc<double>((size_t)GRAM_MAX_BLOCKS * 2 * 64 * 64);
            gram_counter = dev_alloc<unsigned>(1);
            MGB_CUDA(cudaMemset(gram_counter, 0, sizeof(unsigned)));
        }
        if (!kq) kq = vec_alloc<TV>((size_t)L[0].nalloc * m, vec_pad(0), ctx.stream);
    }
    int solveBlockCG(TV* xk, double tol, int max_iter, int* flag, double* resmat, double pinv_tol) {
        ensure_work();
        MGB_CHECK(m >= 1 && m <= 64, "blockCG supports up to 64 right-hand sides");
        ensure_krylov(0, false);
        Level<TV>& lv = L[0];
        const long long n = lv.n, nm = n * m;
        ensure_block_ws();
        const Csr<TV>& A = krylov_A();
        const TV* B = lv.b;
        if (norm(nm, B) == 0.0) {
            dev_zero<TV>(ctx, nm, xk);
            *flag = -9;
            return 0;
        }
        if (pinv_tol < 0) {  // package default: eps * size(B,1) of the GLOBAL problem
            long long nglob = lv.sp.dist ? lv.sp.n_global : n;
            pinv_tol = std::numeric_limits<double>::epsilon() * (double)nglob;
        }
        TV *R = kr, *P = kp, *Q = kAp, *Pn = kq;
        if (norm(nm, xk) == 0.0) dev_copy<TV>(ctx, nm, B, R);
        else residual(A, B, xk, R, 1);                       // R = B - A(X)
        TV* Z = precondition(R);
        dev_copy<TV>(ctx, nm, Z, P);
        std::vector<double> nB, nR;
        colnorms(n, B, nB);
        std::vector<zc> PTQ, PTR, QTZ, Pinv, Alpha, Beta;
        *flag = -1;
        int it = 0;
        for (it = 1; it <= max_iter; ++it) {
            apply_A(A, P, Q, 1);                             // Q = A(P)
            gram(n, P, Q, gram_ws + 0);
            gram(n, P, R, gram_ws + 2 * (size_t)m * m);
            read_matrix(gram_ws + 0, PTQ);
            read_matrix(gram_ws + 2 * (size_t)m * m, PTR);
            general_pinv(m, PTQ, pinv_tol, Pinv);
            matmul_small(m, Pinv, PTR, Alpha, 1.0);          // Alpha = pinv(P'Q) (P'R)
            block_axpy(n, P, Alpha, xk, xk);                 // X += P Alpha
            std::vector<zc> nAlpha(Alpha);
            for (auto& v : nAlpha) v = -v;
            block_axpy(n, Q, nAlpha, R, R);                  // R -= Q Alpha
            colnorms(n, R, nR);
            double mx = 0.0;
            for (int j = 0; j < m; ++j) {
                resmat[(size_t)(it - 1) * m + j] = nR[j] / nB[j];
                mx = std::max(mx, nR[j] / nB[j]);
            }
            if (mx <= tol) {
                *flag = 0;
                break;
            }
            Z = precondition(R);
            gram(n, Q, Z, gram_ws + 4 * (size_t)m * m);
            read_matrix(gram_ws + 4 * (size_t)m * m, QTZ);
            matmul_small(m, Pinv, QTZ, Beta, -1.0);          // Beta = -pinv(P'Q) (Q'Z)
            block_axpy(n, P, Beta, Z, Pn);                   // P = Z + P Beta
            std::swap(P, Pn);
        }
        if (it > max_iter) it = max_iter;
        return it;
    }

    // ---- block QR of an n x m block on the device: W = Q * Betta, Q'Q = I, Betta upper triangular ---------------
    // Cholesky QR applied twice (Gram matrix on the device, m x m Cholesky on the host, Q = W R^{-1} on the device;
    // the second pass restores orthogonality to working precision).  If the Gram matrix is not numerically
    // positive definite the first pass is shifted (shifted CholeskyQR3).  KrylovMethods uses LAPACK's Householder
    // qr!; the factors differ by unit-modulus column scalings only, which the block-Hessenberg least-squares
    // residual does not see.  Wio holds the block on entry and Q on return; tmp is a spare block.
    void block_qr(long long n, TV*& Wio, TV*& tmp, std::vector<zc>& Betta) {
        std::vector<zc> G, R, Rinv, Racc;
        Racc.assign((size_t)m * m, zc(0, 0));
        for (int i = 0; i < m; ++i) Racc[(size_t)i * m + i] = 1.0;
        int passes = 2;
        for (int pass = 0; pass < passes; ++pass) {
            gram(n, Wio, Wio, gram_ws);
            read_matrix(gram_ws, G);
            for (int i = 0; i < m; ++i)     // exactly Hermitian
                for (int j = i + 1; j < m; ++j) {
                    const zc a = 0.5 * (G[(size_t)i * m + j] + std::conj(G[(size_t)j * m + i]));
                    G[(size_t)i * m + j] = a;
                    G[(size_t)j * m + i] = std::conj(a);
                }
            if (!cholesky_upper(m, G, R)) {
                MGB_CHECK(pass == 0 && passes == 2, "block QR: Gram matrix stays indefinite after the shifted pass");
                double tr = 0.0;
                for (int i = 0; i < m; ++i) tr += G[(size_t)i * m + i].real();
                long long nglob = L[0].sp.dist ? L[0].sp.n_global : n;
                const double shift = 11.0 * ((double)m * nglob + (double)m * (m + 1)) * 2.220446049250313e-16 * tr;
                for (int i = 0; i < m; ++i) G[(size_t)i * m + i] += std::max(shift, 1e-300);
                MGB_CHECK(cholesky_upper(m, G, R), "block QR: zero block (exact breakdown)");
                passes = 3;
            }
            upper_inverse(m, R, Rinv);
            block_axpy(n, Wio, Rinv, nullptr, tmp);      // tmp = W R^{-1}
            std::swap(Wio, tmp);
            std::vector<zc> Rn;
            matmul_small(m, R, Racc, Rn, 1.0);           // Betta = R_k ... R_1
            Racc.swap(Rn);
        }
        Betta = Racc;
    }

    // ---- KrylovMethods.blockFGMRES with M = one cycle (solveGMRES_MG with nrhs > 1, SolveFuncs.jl:130) ----------
    // Block Arnoldi: one n x m block per inner step, classical block Gram-Schmidt against the whole basis, QR of the
    // new block, block-Hessenberg least squares on the host every inner step; Frobenius norms; max_iter counts
    // restarts.  resvec: restrt * max_iter doubles.
    int solveBlockFGMRES(TV* xk, int restrt, bool flexible, double tol, int max_iter, int* flag, double* resvec,
                         int* nres) {
        ensure_work();
        MGB_CHECK(m >= 1 && m <= 64, "blockFGMRES supports up to 64 right-hand sides");
        Level<TV>& lv = L[0];
        long long nglob = lv.sp.dist ? lv.sp.n_global : lv.n;
        restrt = (int)std::min<long long>(restrt, nglob - 1);
        MGB_CHECK(restrt >= 1 && restrt <= MAXK, "blockFGMRES: restart length must be in 1..32");
        ensure_krylov(restrt, true);
        ensure_block_ws();
        const long long n = lv.n, nm = n * m;
        const size_t ldv = (size_t)lv.nalloc * m;     // one basis block
        const Csr<TV>& A = krylov_A();
        const TV* B = lv.b;
        TV *R = kr, *W = kw, *T = kAp, *Tmp = kq;
        *nres = 0;
        const double rnorm0 = norm(nm, B);
        if (rnorm0 == 0.0) {
            dev_zero<TV>(ctx, nm, xk);
            *flag = -9;
            return 0;
        }
        residual(A, B, xk, R, 1);
        double err = norm(nm, R) / rnorm0;
        if (err < tol) {
            *flag = 0;
            resvec[0] = err;
            *nres = 1;
            return 0;
        }
        *flag = -1;
        int counter = 0, it = 0;
        const int hr = (restrt + 1) * m, hc = restrt * m;
        std::vector<zc> Betta, Tm, Y;
        while (it < max_iter) {
            it += 1;
            std::vector<zc> H((size_t)hr * hc, zc(0, 0)), xi((size_t)hr * m, zc(0, 0));
            dev_copy<TV>(ctx, nm, R, W);
            block_qr(n, W, Tmp, Betta);
            for (int a = 0; a < m; ++a)
                for (int b = 0; b < m; ++b) xi[(size_t)a * m + b] = Betta[(size_t)a * m + b];
            int jdone = 0;
            for (int j = 0; j < restrt; ++j) {
                dev_copy<TV>(ctx, nm, W, kV + (size_t)j * ldv);
                TV* Z = precondition(W);
                if (flexible) dev_copy<TV>(ctx, nm, Z, kZ + (size_t)j * ldv);
                apply_A(A, Z, W, 1);
                counter += 1;
                // T = V'W block by block, then W -= V T (classical block Gram-Schmidt: all of T from the same W)
                std::vector<std::vector<zc>> Ts(j + 1);
                for (int i = 0; i <= j; ++i) {
                    gram(n, kV + (size_t)i * ldv, W, gram_ws);
                    read_matrix(gram_ws, Ts[i]);
                    for (int a = 0; a < m; ++a)
                        for (int b = 0; b < m; ++b) H[(size_t)(i * m + a) * hc + (j * m + b)] = Ts[i][(size_t)a * m + b];
                }
                for (int i = 0; i <= j; ++i) {
                    for (auto& v : Ts[i]) v = -v;
                    block_axpy(n, kV + (size_t)i * ldv, Ts[i], W, W);
                }
                block_qr(n, W, Tmp, Betta);
                for (int a = 0; a < m; ++a)
                    for (int b = 0; b < m; ++b) H[(size_t)((j + 1) * m + a) * hc + (j * m + b)] = Betta[(size_t)a * m + b];
                const int rows = (j + 2) * m, cols = (j + 1) * m;
                std::vector<zc> Hj((size_t)rows * cols), xj((size_t)rows * m);
                for (int a = 0; a < rows; ++a) {
                    for (int b = 0; b < cols; ++b) Hj[(size_t)a * cols + b] = H[(size_t)a * hc + b];
                    for (int b = 0; b < m; ++b) xj[(size_t)a * m + b] = xi[(size_t)a * m + b];
                }
                err = dense_lsq(rows, cols, m, Hj, xj, Y) / rnorm0;
                resvec[counter - 1] = err;
                jdone = j + 1;
                if (err <= tol) {
                    *flag = 0;
                    break;
                }
            }
            // X += Z y (flexible) or X += M(V y); Y holds the least-squares solution of the last inner step
            std::vector<zc> Yi((size_t)m * m);
            if (!flexible) dev_zero<TV>(ctx, nm, T);
            for (int i = 0; i < jdone; ++i) {
                for (int a = 0; a < m; ++a)
                    for (int b = 0; b < m; ++b) Yi[(size_t)a * m + b] = Y[(size_t)(i * m + a) * m + b];
                if (flexible) block_axpy(n, kZ + (size_t)i * ldv, Yi, xk, xk);
                else block_axpy(n, kV + (size_t)i * ldv, Yi, T, T);
            }
            if (!flexible) {
                TV* z = precondition(T);
                dev_axpby<TV>(ctx, nm, VT<TV>::one(), z, VT<TV>::one(), xk, false);
            }
            if (*flag == 0) break;
            if (it < max_iter) residual(A, B, xk, R, 1);
        }
        // W and Tmp may have traded places inside block_qr: keep the members consistent with the buffers they own
        kw = W;
        kq = Tmp;
        *nres = counter;
        return it;
    }

    // ---- KrylovMethods.blockBiCGSTB with M1 = one cycle (solveBiCGSTAB_MG with nrhs > 1, SolveFuncs.jl:95) ------
    // Block BiCGStab of El Guennouni, Jbilou and Sadok in the shape of solveBiCGSTAB above: m x m systems with
    // R~'V instead of the scalar divisions, omega from Frobenius inner products.  resvec: max_iter + 1 doubles.
    int solveBlockBiCGSTAB(TV* xk, double tol, int max_iter, int* flag, double* resvec, int* nprec) {
        ensure_work();
        MGB_CHECK(m >= 1 && m <= 64, "blockBiCGSTB supports up to 64 right-hand sides");
        ensure_krylov(1, true);
        ensure_block_ws();
        Level<TV>& lv = L[0];
        const long long n = lv.n, nm = n * m;
        const Csr<TV>& A = krylov_A();
        const TV* B = lv.b;
        TV *r = kr, *rt = kw, *p = kp, *v = kAp, *phat = kV, *t = kZ, *tmp = kq;
        *nprec = 0;
        const double bnrm2 = norm(nm, B);
        if (bnrm2 == 0.0) {
            dev_zero<TV>(ctx, nm, xk);
            *flag = -9;
            resvec[0] = 0.0;
            return 0;
        }
        residual(A, B, xk, r, 1);
        double err = norm(nm, r) / bnrm2;
        resvec[0] = err;
        if (err < tol) {
            *flag = 0;
            return 0;
        }
        dev_copy<TV>(ctx, nm, r, rt);
        dev_copy<TV>(ctx, nm, r, p);
        std::vector<zc> G, RtR, RtT, Alpha, Beta, nA;
        *flag = -1;
        int it = 0;
        for (it = 1; it <= max_iter; ++it) {
            TV* z = precondition(p);                         // P_hat = M2(M1(P)), M2 = identity
            *nprec += m;
            dev_copy<TV>(ctx, nm, z, phat);
            apply_A(A, phat, v, 1);                          // V = A(P_hat)
            gram(n, rt, v, gram_ws);
            read_matrix(gram_ws, G);
            gram(n, rt, r, gram_ws);
            read_matrix(gram_ws, RtR);
            if (!lu_solve_small(m, m, G, RtR, Alpha)) {      // (R~'V) alpha = R~'R
                *flag = -2;
                break;
            }
            nA = Alpha;
            for (auto& a : nA) a = -a;
            block_axpy(n, v, nA, r, r);                      // S = R - V alpha (in r)
            const double snorm = norm(nm, r) / bnrm2;
            if (snorm < tol) {
                block_axpy(n, phat, Alpha, xk, xk);          // X += P_hat alpha
                resvec[it] = snorm;
                *flag = -3;
                return it - 1;
            }
            z = precondition(r);                             // S_hat
            *nprec += m;
            apply_A(A, z, t, 1);                             // T = A(S_hat)
            const zc ts = dot(nm, t, r), tt = dot(nm, t, t);
            const zc omega = ts / tt;
            block_axpy(n, phat, Alpha, xk, xk);              // X += P_hat alpha + omega S_hat
            dev_axpby<TV>(ctx, nm, to_tv(omega), z, VT<TV>::one(), xk, false);
            dev_axpby<TV>(ctx, nm, to_tv(-omega), t, VT<TV>::one(), r, false);   // R = S - omega T
            err = norm(nm, r) / bnrm2;
            resvec[it] = err;
            if (err <= tol) {
                *flag = 0;
                break;
            }
            if (omega == zc(0.0, 0.0)) {
                *flag = -2;
                break;
            }
            gram(n, rt, t, gram_ws);
            read_matrix(gram_ws, RtT);
            for (auto& a : RtT) a = -a;
            if (!lu_solve_small(m, m, G, RtT, Beta)) {       // (R~'V) beta = -R~'T
                *flag = -2;
                break;
            }
            dev_axpby<TV>(ctx, nm, to_tv(-omega), v, VT<TV>::one(), p, false);   // P - omega V (in p)
            block_axpy(n, p, Beta, r, tmp);                  // P = R + (P - omega V) beta
            std::swap(p, tmp);
        }
        kp = p;
        kq = tmp;
        if (it > max_iter) it = max_iter;
        return it;
    }

    // ---- KrylovMethods.bicgstb with M1 = one cycle, M2 = identity (solveBiCGSTAB_MG, SolveFuncs.jl:85-99) ----
    // van der Vorst's preconditioned BiCGStab as in the "Templates" book, which the package follows.
    // resvec: max_iter + 1 doubles (resvec[0] = initial relative residual).  flag: 0 converged, -1 max_iter,
    // -2 breakdown (rho == 0 or omega == 0), -3 converged after the first half of an iteration, -9 b == 0.
    // Returns the number of completed iterations; *nprec = cycles applied (2 per iteration, SolveFuncs.jl:97).
    int solveBiCGSTAB(TV* xk, double tol, int max_iter, int* flag, double* resvec, int* nprec) {
        ensure_work();
        if (m > 1) return solveBlockBiCGSTAB(xk, tol, max_iter, flag, resvec, nprec);
        ensure_krylov(1, true);
        Level<TV>& lv = L[0];
        const long long n = lv.n;
        const Csr<TV>& A = krylov_A();
        const TV* b = lv.b;
        TV *r = kr, *rt = kw, *p = kp, *v = kAp, *phat = kV, *t = kZ;
        *nprec = 0;
        const double bnrm2 = norm(n, b);
        if (bnrm2 == 0.0) {
            dev_zero<TV>(ctx, n, xk);
            *flag = -9;
            resvec[0] = 0.0;
            return 0;
        }
        residual(A, b, xk, r, 1);
        double err = norm(n, r) / bnrm2;
        resvec[0] = err;
        if (err < tol) {
            *flag = 0;
            return 0;
        }
        dev_copy<TV>(ctx, n, r, rt);
        zc omega(1.0, 0.0), alpha(1.0, 0.0), rho1(1.0, 0.0);
        *flag = -1;
        int it = 0;
        for (it = 1; it <= max_iter; ++it) {
            const zc rho = dot(n, rt, r);
            if (rho == zc(0.0, 0.0)) {
                *flag = -2;
                break;
            }
            if (it > 1) {
                const zc beta = (rho / rho1) * (alpha / omega);
                Launch La(ctx, K_VECTOR, 0, 4.0 * n * sizeof(TV));
                bicg_p_kernel<TV><<<ctx.ew_blocks(n), 256, 0, ctx.stream>>>(n, to_tv(beta), to_tv(omega), r, v, p);
                MGB_LAUNCH_CHECK();
            } else {
                dev_copy<TV>(ctx, n, r, p);
            }
            TV* z = precondition(p);                         // p_hat = M2(M1(p)), M2 = identity
            *nprec += 1;
            dev_copy<TV>(ctx, n, z, phat);
            apply_A(A, phat, v, 1);                          // v = A(p_hat)
            alpha = rho / dot(n, rt, v);
            dev_axpby<TV>(ctx, n, to_tv(-alpha), v, VT<TV>::one(), r, false);   // s = r - alpha*v (in r)
            const double snorm = norm(n, r) / bnrm2;
            if (snorm < tol) {
                Launch La(ctx, K_VECTOR, 0, 3.0 * n * sizeof(TV));
                bicg_x_kernel<TV><<<ctx.ew_blocks(n), 256, 0, ctx.stream>>>(n, to_tv(alpha), phat, VT<TV>::zero(), phat, xk, 0);
                MGB_LAUNCH_CHECK();
                resvec[it] = snorm;
                *flag = -3;
                return it - 1;
            }
            z = precondition(r);                             // s_hat
            *nprec += 1;
            apply_A(A, z, t, 1);                             // t = A(s_hat)
            const zc ts = dot(n, t, r), tt = dot(n, t, t);
            omega = ts / tt;
            {
                Launch La(ctx, K_VECTOR, 0, 4.0 * n * sizeof(TV));
                bicg_x_kernel<TV><<<ctx.ew_blocks(n), 256, 0, ctx.stream>>>(n, to_tv(alpha), phat, to_tv(omega), z, xk, 1);
                MGB_LAUNCH_CHECK();
            }
            dev_axpby<TV>(ctx, n, to_tv(-omega), t, VT<TV>::one(), r, false);   // r = s - omega*t
            err = norm(n, r) / bnrm2;
            resvec[it] = err;
            if (err <= tol) {
                *flag = 0;
                break;
            }
            if (omega == zc(0.0, 0.0)) {
                *flag = -2;
                break;
            }
            rho1 = rho;
        }
        if (it > max_iter) it = max_iter;
        return it;
    }

    // ---- KrylovMethods.fgmres (solveGMRES_MG, SolveFuncs.jl:120-132; coarsest "GMRES", MGcycle.jl:152-168) ----
    // Restarted right-preconditioned (F)GMRES on level `level` (1-based) with workspace ws; prec(v) returns a
    // buffer holding M v.  max_iter counts restarts.
    struct FgmresWs {
        TV *r = nullptr, *w = nullptr, *t = nullptr, *V = nullptr, *Z = nullptr;
        long long ld = 0;
    };
    int fgmres_core(const Csr<TV>& A, int level, long long n, const TV* b, TV* xk, int restrt, bool flexible,
                    double tol, int max_iter, const PrecFn& prec, const FgmresWs& ws, int* flag, double* resvec,
                    int* nres) {
        const int l = level - 1;
        restrt = (int)std::min<long long>(restrt, n - 1);
        MGB_CHECK(restrt >= 1 && restrt <= MAXK, "fgmres: restart length must be in 1..32");
        MGB_CHECK(!flexible || ws.Z, "fgmres: flexible variant needs the Z workspace");
        *nres = 0;
        const double rnorm0 = norm(n, b, l);
        if (rnorm0 == 0.0) {
            dev_zero<TV>(ctx, n, xk);
            *flag = -9;
            return 0;
        }
        residual(A, b, xk, ws.r, level);
        double err = norm(n, ws.r, l) / rnorm0;
        if (err < tol) {
            *flag = 0;
            if (resvec) resvec[0] = err;
            *nres = 1;
            return 0;
        }
        *flag = -1;
        int counter = 0, it = 0;
        const int ldh = restrt;
        const long long ldv = ws.ld;
        TV* kw_ = ws.w;
        while (it < max_iter) {
            it += 1;
            std::vector<zc> H((size_t)(restrt + 1) * ldh, zc(0, 0)), xi(restrt + 1, zc(0, 0)), y;
            double betta = norm(n, ws.r, l);
            xi[0] = betta;
            dev_axpby<TV>(ctx, n, VT<TV>::from_real(1.0 / betta), ws.r, VT<TV>::zero(), kw_, true);  // w = r/betta
            dev_copy<TV>(ctx, n, kw_, ws.V);
            int jdone = 0;
            for (int j = 0; j < restrt; ++j) {
                TV* z = prec(kw_);                                               // z = M(w)
                if (flexible) dev_copy<TV>(ctx, n, z, ws.Z + (size_t)j * ldv);
                apply_A(A, z, kw_, level);                                       // w = A(z)
                counter += 1;
                dev_multi_dot<TV>(ctx, n, ws.V, ldv, j + 1, kw_, ctx.scal + 16);  // t = V'w
                allreduce(l, ctx.scal + 16, 2 * (j + 1));
                std::vector<double> hv(2 * (j + 1));
                read_scalars(ctx, ctx.scal + 16, 2 * (j + 1), hv.data());
                std::vector<TV> coef(j + 1);
                for (int i = 0; i <= j; ++i) {
                    H[(size_t)i * ldh + j] = zc(hv[2 * i], hv[2 * i + 1]);
                    coef[i] = to_tv(-H[(size_t)i * ldh + j]);
                }
                dev_multi_axpy<TV>(ctx, n, ws.V, ldv, j + 1, coef.data(), 1.0, kw_, ctx.scal + 12);  // w -= V t, ||w||^2
                allreduce(l, ctx.scal + 12, 1);
                double nw2;
                read_scalars(ctx, ctx.scal + 12, 1, &nw2);
                betta = std::sqrt(nw2);
                H[(size_t)(j + 1) * ldh + j] = betta;
                dev_axpby<TV>(ctx, n, VT<TV>::from_real(1.0 / betta), kw_, VT<TV>::zero(), kw_, true);  // w *= 1/betta
                if (j + 1 < restrt) dev_copy<TV>(ctx, n, kw_, ws.V + (size_t)(j + 1) * ldv);
                err = hessenberg_lsq(H, ldh, j + 2, j + 1, xi, y) / rnorm0;
                if (resvec) resvec[counter - 1] = err;
                jdone = j + 1;
                if (err <= tol) {
                    *flag = 0;
                    break;
                }
            }
            hessenberg_lsq(H, ldh, jdone + 1, jdone, xi, y);                      // y = pinv(H)*xi
            std::vector<TV> coef(jdone);
            for (int i = 0; i < jdone; ++i) coef[i] = to_tv(y[i]);
            if (flexible) {
                dev_multi_axpy<TV>(ctx, n, ws.Z, ldv, jdone, coef.data(), 1.0, xk, nullptr);  // x += Z y
            } else {
                dev_multi_axpy<TV>(ctx, n, ws.V, ldv, jdone, coef.data(), 0.0, ws.t, nullptr);  // V y
                TV* z = prec(ws.t);
                dev_axpby<TV>(ctx, n, VT<TV>::one(), z, VT<TV>::one(), xk, false);        // x += M(V y)
            }
            if (*flag == 0) break;
            if (it < max_iter) residual(A, b, xk, ws.r, level);
        }
        *nres = counter;
        return it;
    }
    int solveFGMRES(TV* xk, int restrt, bool flexible, double tol, int max_iter, int* flag, double* resvec,
                    int* nres) {
        ensure_work();
        if (m > 1) return solveBlockFGMRES(xk, restrt, flexible, tol, max_iter, flag, resvec, nres);
        Level<TV>& lv = L[0];
        restrt = (int)std::min<long long>(restrt, lv.n - 1);
        MGB_CHECK(restrt >= 1 && restrt <= MAXK, "fgmres: restart length must be in 1..32");
        ensure_krylov(restrt, true);
        FgmresWs ws;
        ws.r = kr; ws.w = kw; ws.t = kAp; ws.V = kV; ws.Z = kZ;
        ws.ld = lv.nalloc;   // basis columns are allocated with ghost space
        PrecFn prec = [this](const TV* v) -> TV* { return precondition(v); };
        return fgmres_core(krylov_A(), 1, lv.n, lv.b, xk, restrt, flexible, tol, max_iter, prec, ws, flag, resvec, nres);
    }

    // ---- host <-> device layout helpers -------------------------------------------------------
    void h2d_vec(const void* host, TV* dst, long long n) {
        const long long nm = n * m;
        if (m == 1) {
            MGB_CUDA(cudaMemcpyAsync(dst, host, nm * sizeof(TV), cudaMemcpyHostToDevice, ctx.stream));
            return;
        }
        if (hstage_n < nm) {
            dev_free(hstage);
            hstage = dev_alloc<TV>(nm);
            hstage_n = nm;
        }
        MGB_CUDA(cudaMemcpyAsync(hstage, host, nm * sizeof(TV), cudaMemcpyHostToDevice, ctx.stream));
        colmajor_to_rhsfast_kernel<TV><<<ctx.ew_blocks(nm), 256, 0, ctx.stream>>>(n, m, hstage, dst);
        MGB_LAUNCH_CHECK();
    }
    void d2h_vec(const TV* src, void* host, long long n) {
        const long long nm = n * m;
        if (m == 1) {
            MGB_CUDA(cudaMemcpyAsync(host, src, nm * sizeof(TV), cudaMemcpyDeviceToHost, ctx.stream));
            ctx.sync();
            return;
        }
        if (hstage_n < nm) {
            dev_free(hstage);
            hstage = dev_alloc<TV>(nm);
            hstage_n = nm;
        }
        rhsfast_to_colmajor_kernel<TV><<<ctx.ew_blocks(nm), 256, 0, ctx.stream>>>(n, m, src, hstage);
        MGB_LAUNCH_CHECK();
        MGB_CUDA(cudaMemcpyAsync(host, hstage, nm * sizeof(TV), cudaMemcpyDeviceToHost, ctx.stream));
        ctx.sync();
    }
};

}  // namespace mgb200
