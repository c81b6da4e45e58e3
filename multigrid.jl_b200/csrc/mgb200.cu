// C ABI of libmgb200.so (declared in include/mgb200.h).
#include "../../include/mgb200.h"

#include <cmath>
#include <array>
#include <cstring>
#include <map>
#include <string>
#include <type_traits>

#include <cuda_profiler_api.h>

#include "solver.cuh"

using namespace mgb200;

struct mgb200_hierarchy {
    int val_type;
    HierarchyBase* impl;
};

static thread_local std::string g_last_error;

#define MGB_TRY try {
#define MGB_CATCH                                         \
    }                                                     \
    catch (const Error& e) {                              \
        g_last_error = e.what();                          \
        return e.code;                                    \
    }                                                     \
    catch (const std::exception& e) {                     \
        g_last_error = e.what();                          \
        return -1;                                        \
    }                                                     \
    return 0;

template <typename F>
static void dispatch(mgb200_handle h, F&& f) {
    MGB_CHECK(h && h->impl, "null handle");
    if (h->val_type == MGB200_FP64) {
        auto* H = static_cast<Hierarchy<double>*>(h->impl);
        MGB_CUDA(cudaSetDevice(H->ctx.device));
        f(H);
    } else if (h->val_type == MGB200_CFP64) {
        auto* H = static_cast<Hierarchy<cplx>*>(h->impl);
        MGB_CUDA(cudaSetDevice(H->ctx.device));
        f(H);
    } else if (h->val_type == MGB200_FP32) {
        auto* H = static_cast<Hierarchy<float>*>(h->impl);
        MGB_CUDA(cudaSetDevice(H->ctx.device));
        f(H);
    } else {
        auto* H = static_cast<Hierarchy<cplxf>*>(h->impl);
        MGB_CUDA(cudaSetDevice(H->ctx.device));
        f(H);
    }
}
// generic body for all value types; `H` is the typed hierarchy pointer
#define MGB_BOTH(h, ...) dispatch(h, [&](auto* H) { __VA_ARGS__; })

template <typename TV>
static void spmatmul_impl(Hierarchy<TV>* H, int level, int which, double alpha, const void* x, double beta,
                          void* y) {
    MGB_CHECK(level >= 1 && level <= H->levels, "level out of range");
    MGB_CHECK(which >= 0 && which <= 2, "which must be 0 (A), 1 (P) or 2 (R)");
    Level<TV>& lv = H->L[level - 1];
    MGB_CHECK(!lv.sp.dist, "mgb200_spmatmul is not available on row-partitioned levels");
    int mode;
    if (alpha == 1.0 && beta == 0.0) mode = MODE_SPMV;
    else if (alpha == 1.0 && beta == 1.0) mode = MODE_ADD;
    else if (alpha == -1.0 && (beta == 1.0 || beta == 0.0)) mode = MODE_RESID;
    else throw Error(-1, "mgb200_spmatmul: (alpha,beta) must be one of (1,0),(1,1),(-1,1),(-1,0)");
    long long nx, ny;
    const int m = H->m;
    Context& ctx = H->ctx;
    auto run = [&](auto& M) {
        nx = M.n_cols;
        ny = M.n_rows;
        TV* dx = vec_alloc<TV>(nx * m, 0, ctx.stream);
        TV* dy = vec_alloc<TV>(ny * m, 0, ctx.stream);
        TV* db = vec_alloc<TV>(ny * m, 0, ctx.stream);
        H->h2d_vec(x, dx, nx);
        if (beta == 1.0) H->h2d_vec(y, dy, ny);
        if (mode == MODE_RESID) {
            if (beta == 1.0) MGB_CUDA(cudaMemcpyAsync(db, dy, ny * m * sizeof(TV), cudaMemcpyDeviceToDevice, ctx.stream));
            else MGB_CUDA(cudaMemsetAsync(db, 0, ny * m * sizeof(TV), ctx.stream));
        }
        typedef typename std::remove_reference<decltype(*M.val)>::type TA;
        csr_apply<TA, TV>(ctx, M, mode, dx, db, nullptr, dy, m, K_SPMV, level);
        H->d2h_vec(dy, y, ny);
        dev_free(dx);
        dev_free(dy);
        dev_free(db);
    };
    if (which == 0) {
        MGB_CHECK(lv.A.present(), "matrix not uploaded");
        run(lv.A);
    } else if (which == 1) {
        MGB_CHECK(lv.P.present(), "matrix not uploaded");
        run(lv.P);
    } else {
        MGB_CHECK(lv.R.present(), "matrix not uploaded");
        run(lv.R);
    }
}

// outer (double-precision, Krylov only) handle over a single-precision hierarchy
template <typename TD, typename TS>
static HierarchyBase* make_mixed(Hierarchy<TS>* in) {
    MGB_CUDA(cudaSetDevice(in->ctx.device));
    MGB_CHECK(!in->comm.active(), "mixed precision over a row-partitioned hierarchy is not supported");
    MGB_CHECK(in->L[0].n > 0, "upload the single-precision hierarchy first");
    const int64_t one[1] = {1};
    auto* out = new Hierarchy<TD>(1, in->m, in->cycle_type, 0, one, one, in->ctx.device);
    out->krylov_only = true;
    out->L[0].n = in->L[0].n;
    out->L[0].nalloc = in->L[0].n;
    for (auto& e : out->mix_ev) MGB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    out->ext_prec = [out, in](const TD* r) -> TD* {
        MGB_CHECK(in->m == out->m, "nrhs of the two mixed-precision handles differ (mgb200_adjust_nrhs on both)");
        in->ensure_work();
        const long long nm = in->L[0].n * in->m;
        cudaStream_t si = in->ctx.stream, so = out->ctx.stream;
        MGB_CUDA(cudaEventRecord(out->mix_ev[0], so));
        MGB_CUDA(cudaStreamWaitEvent(si, out->mix_ev[0], 0));
        const long long l0 = in->ctx.launches;
        convert_kernel<TD, TS><<<in->ctx.ew_blocks(nm), 256, 0, si>>>(nm, r, in->L[0].b);      // bl[:] .= b
        MGB_LAUNCH_CHECK();
        TS* z = in->precondition(in->L[0].b);                                                   // z .= 0; cycle
        convert_kernel<TS, TD><<<in->ctx.ew_blocks(nm), 256, 0, si>>>(nm, z, out->L[0].x0);    // z2[:] .= z
        MGB_LAUNCH_CHECK();
        MGB_CUDA(cudaEventRecord(out->mix_ev[1], si));
        MGB_CUDA(cudaStreamWaitEvent(so, out->mix_ev[1], 0));
        out->ctx.launches += 2 + (in->ctx.launches - l0);
        return out->L[0].x0;
    };
    return out;
}

template <typename TA>
static void host_tma_plan_impl(const HostPatterns<double>& hp, int tile, int64_t* info, int32_t* win, int32_t* delta,
                               int32_t* soff) {
    // the plan depends on the offsets and on the element size only: re-type the dictionary
    HostPatterns<TA> h;
    h.ok = hp.ok;
    h.rowrel = hp.rowrel;
    h.pat_off = hp.pat_off;
    h.delta = hp.delta;
    h.val.resize(hp.val.size());
    TmaPlan P;
    std::vector<int> so;
    if (!build_tma_plan<TA>(h, tile, P, so)) return;
    info[0] = 1;
    info[1] = P.nwin;
    info[2] = P.total;
    info[3] = P.centre;
    for (int g = 0; g < P.nwin; ++g) {
        win[3 * g] = P.w[g].lo_even;
        win[3 * g + 1] = P.w[g].len;
        win[3 * g + 2] = P.w[g].sbase;
    }
    for (size_t k = 0; k < so.size(); ++k) {
        delta[k] = hp.delta[k];
        soff[k] = so[k];
    }
}

// CPU replay of a box_kernel launch (box.cuh): the tile plan, the copies and the per-thread function are the
// kernel's own __host__ __device__ code; the stages are host buffers that persist from tile to tile like shared memory.
template <int SHAPE, int MODE, bool DPAT, int RZ>
static void host_box_run(const BoxPlan& P, const BoxCoef<double>& C0, const std::vector<double>& ctab, const std::vector<double>& dtab,
                         int NB, int ctas, const uint16_t* pid, const double* x, const double* b, const double* d, double* y,
                         long long* fast_rows) {
    constexpr bool NEED_B = (MODE == 2 || MODE == 3), NEED_D = (MODE == 3 && !DPAT), NEED_PW = (MODE == 4);
    const size_t stage_bytes = box_stage_bytes<double>(P, RZ, NB, NEED_B, NEED_D, NEED_PW);
    const int bcap = NB + 4, pcap = NB + 16;
    const size_t rb = box_rec_bytes(RZ);
    std::vector<unsigned char> rec(rb);
    for (int cta = 0; cta < ctas; ++cta) {
        std::vector<unsigned char> stage(stage_bytes, 0);
        for (long long tile = cta; tile < P.ntiles; tile += ctas) {
            box_plan_tile<double>(P, RZ, NB, (int)tile, rec.data());
            const BoxCopy* cp = reinterpret_cast<const BoxCopy*>(rec.data() + BOX_DESC_BYTES);
            std::memcpy(stage.data(), rec.data(), BOX_DESC_BYTES);
            for (int i = 0; i < box_ncopies(RZ); ++i) {
                const BoxCopy& C = cp[i];
                if (!C.bytes || (C.what == 1 && !NEED_B) || (C.what == 2 && !NEED_D) || (C.what == 5 && !NEED_PW)) continue;
                MGB_CHECK((size_t)C.dst + C.bytes <= stage_bytes && C.dst % 16 == 0 && C.bytes % 16 == 0, "copy outside the stage or unaligned");
                const unsigned char* src = C.what == 0 ? reinterpret_cast<const unsigned char*>(x + C.src)
                                           : (C.what == 1 ? reinterpret_cast<const unsigned char*>(b + C.src)
                                                          : (C.what == 2 ? reinterpret_cast<const unsigned char*>(d + C.src)
                                                                         : reinterpret_cast<const unsigned char*>(pid + C.src)));
                MGB_CHECK(((C.what == 3 || C.what == 5) ? C.src % 8 : C.src % 2) == 0, "unaligned source of a bulk copy");
                std::memcpy(stage.data() + C.dst, src, C.bytes);
            }
            const BoxDesc& T = *reinterpret_cast<const BoxDesc*>(stage.data());
            const double* sx = reinterpret_cast<const double*>(stage.data() + BOX_DESC_BYTES);
            const uint16_t* sp = reinterpret_cast<const uint16_t*>(sx + P.xtotal);
            const double* sb = reinterpret_cast<const double*>(sp + RZ * pcap);
            const double* sd = sb + RZ * bcap;
            for (int t = 0; t < T.nb; ++t) {
                const double* xc[RZ + 2];
                const uint16_t* pwc[RZ + 2];
                const double* bp[RZ];
                const double* dp[RZ];
                int pat[RZ];
                bool fast = true;
                for (int w = 0; w < RZ + 2; ++w) {
                    xc[w] = sx + T.xoff[w] + t;
                    pwc[w] = sp + RZ * pcap + (NEED_PW ? T.pwoff[w] : 0) + t;
                }
                for (int j = 0; j < RZ; ++j) {
                    bp[j] = sb + T.boff[j] + t;
                    dp[j] = sd + T.boff[j] + t;
                    pat[j] = j < T.nrp ? (int)sp[T.poff[j] + t] : P.p0;
                    fast = fast && pat[j] == P.p0;
                }
                double out[RZ];
                if (fast) box_thread<double, SHAPE, MODE, DPAT, RZ, true>(C0, ctab.data(), dtab.data(), P.NP, P.S, xc, bp, dp, pat, out, pwc);
                else box_thread<double, SHAPE, MODE, DPAT, RZ, false>(C0, ctab.data(), dtab.data(), P.NP, P.S, xc, bp, dp, pat, out, pwc);
                if (fast) *fast_rows += T.nrp;
                for (int j = 0; j < T.nrp; ++j) y[T.r0 + t + (long long)j * P.S2] = out[j];
            }
        }
    }
}

// the one-row-per-thread dictionary walk of pat_kernel on the CPU: the reference every other kernel form must match bit for bit
template <int MODE, bool DPAT>
static void host_pattern_walk(const HostPatterns<double>& hp, long long S2, long long n_rows, const std::vector<PatEntry<double>>& ent,
                              const double* dpat, const double* x, const double* b, const double* d, double* y) {
    const LinesView<double> Vg = lines_global_view<double>(S2, hp.pid.data(), x, b, d);
    for (long long row = 0; row < n_rows; ++row)
        y[row] = pat_row_walk<double, double, MODE, DPAT>(S2, Vg, row, hp.pat_off.data(), ent.data(), dpat);
}

// CPU replay of the grid-hinted transfer kernels (grid_xfer.cuh): their per-row functions, row by row
static void host_gx_run(const GridXfer& X, const std::vector<double>& tab, const double* x, double* y) {
    if (X.kind == 1) {
        const long long cs2 = (long long)X.N[0] * X.N[1];
        for (int k = 0; k < X.n[2]; ++k)
            for (int j = 0; j < X.n[1]; ++j) {
                const double* q = x + ((long long)(k >> 1) * X.N[1] + (j >> 1)) * X.N[0];
                double* xl = y + ((long long)k * X.n[1] + j) * X.n[0];
                for (int i = 0; i < X.n[0]; ++i) xl[i] = gxp_row<double, double>(tab.data(), q, X.N[0], cs2, i, j & 1, k & 1, xl[i]);
            }
    } else {
        const long long S = X.n[0], S2 = (long long)X.n[0] * X.n[1];
        for (int K = 0; K < X.N[2]; ++K)
            for (int J = 0; J < X.N[1]; ++J) {
                const int cy = gx_class(J, X.N[1]), cz = gx_class(K, X.N[2]);
                for (int I = 0; I < X.N[0]; ++I) {
                    const int cx = gx_class(I, X.N[0]);
                    y[((long long)K * X.N[1] + J) * X.N[0] + I] =
                        gxr_row<double, double>(tab.data() + (cx + 3 * cy + 9 * cz) * 27, x + S2 * (2LL * K) + S * (2LL * J) + 2 * I, S, S2,
                                                gx_allowed(cx, X.N[0]), gx_allowed(cy, X.N[1]), gx_allowed(cz, X.N[2]));
                }
            }
    }
}

#include "multi.cuh"

extern "C" {

const char* mgb200_last_error(void) { return g_last_error.c_str(); }
int mgb200_version(void) { return 100; }

int mgb200_create(mgb200_handle* h, int val_type, int levels, int nrhs, char cycle_type, int relax_kind,
                  const int64_t* relax_pre, const int64_t* relax_post, int device) {
    MGB_TRY
    MGB_CHECK(h != nullptr, "null handle pointer");
    MGB_CHECK(val_type >= MGB200_FP64 && val_type <= MGB200_CFP32, "val_type must be MGB200_FP64, _CFP64, _FP32 or _CFP32");
    MGB_CHECK(relax_kind == 0 || relax_kind == 1, "relax_kind must be 0 (diagonal) or 1 (Jac-GMRES)");
    MGB_CHECK(relax_pre && relax_post, "relax_pre / relax_post required");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        throw Error(-2, std::string("mgb200 needs a CUDA device and has no CPU fallback: ") +
                            (e != cudaSuccess ? cudaGetErrorString(e) : "no device found"));
    MGB_CHECK(device >= 0 && device < ndev, "device index out of range");
    mgb200_hierarchy* hh = new mgb200_hierarchy;
    hh->val_type = val_type;
    hh->impl = nullptr;
    try {
        if (val_type == MGB200_FP64)
            hh->impl = new Hierarchy<double>(levels, nrhs, cycle_type, relax_kind, relax_pre, relax_post, device);
        else if (val_type == MGB200_CFP64)
            hh->impl = new Hierarchy<cplx>(levels, nrhs, cycle_type, relax_kind, relax_pre, relax_post, device);
        else if (val_type == MGB200_FP32)
            hh->impl = new Hierarchy<float>(levels, nrhs, cycle_type, relax_kind, relax_pre, relax_post, device);
        else
            hh->impl = new Hierarchy<cplxf>(levels, nrhs, cycle_type, relax_kind, relax_pre, relax_post, device);
    } catch (...) {
        delete hh;
        throw;
    }
    hh->impl->val_type = val_type;
    *h = hh;
    MGB_CATCH
}

int mgb200_create_mixed(mgb200_handle* outer, mgb200_handle inner) {
    MGB_TRY
    MGB_CHECK(outer && inner && inner->impl, "null handle");
    MGB_CHECK(inner->val_type == MGB200_FP32 || inner->val_type == MGB200_CFP32,
              "mgb200_create_mixed: the inner hierarchy must be MGB200_FP32 or MGB200_CFP32");
    mgb200_hierarchy* hh = new mgb200_hierarchy;
    hh->impl = nullptr;
    try {
        if (inner->val_type == MGB200_FP32) {
            hh->val_type = MGB200_FP64;
            hh->impl = make_mixed<double, float>(static_cast<Hierarchy<float>*>(inner->impl));
        } else {
            hh->val_type = MGB200_CFP64;
            hh->impl = make_mixed<cplx, cplxf>(static_cast<Hierarchy<cplxf>*>(inner->impl));
        }
    } catch (...) {
        delete hh;
        throw;
    }
    hh->impl->val_type = hh->val_type;
    *outer = hh;
    MGB_CATCH
}

int mgb200_destroy(mgb200_handle h) {
    MGB_TRY
    if (h) {
        delete h->impl;
        delete h;
    }
    MGB_CATCH
}

int mgb200_upload_level(mgb200_handle h, int level, int64_t n, int64_t nc, const int64_t* a_colptr,
                        const int64_t* a_rowval, const void* a_nzval, const int64_t* p_colptr,
                        const int64_t* p_rowval, const void* p_nzval, const int64_t* r_colptr,
                        const int64_t* r_rowval, const void* r_nzval, const void* d, int index_base) {
    MGB_TRY
    MGB_BOTH(h, H->upload_level(level, n, nc, a_colptr, a_rowval, a_nzval, p_colptr, p_rowval, p_nzval,
                                r_colptr, r_rowval, r_nzval, d, index_base));
    MGB_CATCH
}

int mgb200_set_level_grid(mgb200_handle h, int level, int dim, const int64_t* n_fine_nodes, const int64_t* n_coarse_nodes) {
    MGB_TRY
    MGB_BOTH(h, H->set_level_grid(level, dim, n_fine_nodes, n_coarse_nodes));
    MGB_CATCH
}

int mgb200_upload_coarsest(mgb200_handle h, int64_t n, const int64_t* colptr, const int64_t* rowval,
                           const void* nzval, int index_base) {
    MGB_TRY
    MGB_BOTH(h, H->upload_coarsest(n, colptr, rowval, nzval, index_base));
    MGB_CATCH
}

int mgb200_replace_matrix(mgb200_handle h, int64_t n, const int64_t* colptr, const int64_t* rowval, const void* nzval,
                          int index_base, int relax_type, const double* relax_param, int* done) {
    MGB_TRY
    MGB_CHECK(colptr && rowval && nzval && done, "null argument");
    *done = 0;
    MGB_BOTH(h, *done = H->replace_matrix(n, colptr, rowval, nzval, index_base, relax_type, relax_param) ? 1 : 0);
    MGB_CATCH
}

int mgb200_download_values(mgb200_handle h, int level, int which, void* nzval, int64_t nnz) {
    MGB_TRY
    MGB_CHECK(nzval && which >= 0 && which <= 2, "bad argument");
    MGB_BOTH(h, H->download_values(level, which, nzval, nnz));
    MGB_CATCH
}

int mgb200_download_relax_prec(mgb200_handle h, int level, void* d) {
    MGB_TRY
    MGB_CHECK(d, "null argument");
    MGB_BOTH(h, H->download_relax_prec(level, d));
    MGB_CATCH
}

int mgb200_upload_coarsest_gmres(mgb200_handle h, int64_t n, const int64_t* colptr, const int64_t* rowval,
                                 const void* nzval, const void* d, int index_base) {
    MGB_TRY
    MGB_CHECK(colptr && rowval && nzval && d, "null argument");
    MGB_BOTH(h, H->upload_coarsest_gmres(n, colptr, rowval, nzval, d, index_base));
    MGB_CATCH
}

int mgb200_set_krylov_matrix(mgb200_handle h, int64_t n, const int64_t* colptr, const int64_t* rowval,
                             const void* nzval, int index_base) {
    MGB_TRY
    MGB_BOTH(h, H->set_krylov_matrix(n, colptr, rowval, nzval, index_base));
    MGB_CATCH
}

int mgb200_dist_unique_id(char* out128) {
    MGB_TRY
    MGB_CHECK(out128, "null output");
    nccl().load();
    ncclUniqueId id;
    MGB_NCCL(nccl().GetUniqueId(&id));
    static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
    std::memcpy(out128, &id, sizeof(id));
    MGB_CATCH
}

int mgb200_dist_init(mgb200_handle h, int rank, int world, const char* unique_id128) {
    MGB_TRY
    MGB_CHECK(world == 1 || unique_id128, "null unique id");
    MGB_BOTH(h, H->dist_init(rank, world, unique_id128));
    MGB_CATCH
}

int mgb200_dist_upload_level(mgb200_handle h, int level, int64_t n_global, const int64_t* row_offsets,
                             int64_t nc_global, const int64_t* coarse_row_offsets, const int64_t* a_colptr,
                             const int64_t* a_rowval, const void* a_nzval, const int64_t* p_colptr,
                             const int64_t* p_rowval, const void* p_nzval, const int64_t* r_colptr,
                             const int64_t* r_rowval, const void* r_nzval, const void* d, int index_base) {
    MGB_TRY
    MGB_CHECK(row_offsets && coarse_row_offsets && a_colptr && a_rowval && a_nzval && p_colptr && p_rowval &&
                  p_nzval && r_colptr && r_rowval && r_nzval && d, "null argument");
    MGB_BOTH(h, H->dist_upload_level(level, n_global, row_offsets, nc_global, coarse_row_offsets, a_colptr,
                                     a_rowval, a_nzval, p_colptr, p_rowval, p_nzval, r_colptr, r_rowval,
                                     r_nzval, d, index_base));
    MGB_CATCH
}

int mgb200_host_plan_ghosts(int64_t nnz, const int64_t* cols, int64_t lo, int64_t hi, int64_t* ghosts,
                            int64_t* n_ghost, int64_t* local_cols) {
    MGB_TRY
    MGB_CHECK(cols && ghosts && n_ghost && local_cols, "null argument");
    std::vector<long long> g;
    collect_ghosts(cols, nnz, lo, hi, g);
    sort_unique(g);
    *n_ghost = (int64_t)g.size();
    for (size_t i = 0; i < g.size(); ++i) ghosts[i] = g[i];
    for (int64_t k = 0; k < nnz; ++k) local_cols[k] = to_local(cols[k], lo, hi, g);
    MGB_CATCH
}

int mgb200_dist_info(mgb200_handle h, int64_t* out) {
    MGB_TRY
    MGB_CHECK(out, "null output");
    MGB_BOTH(h, {
        if (H->comm.active()) H->ensure_work();
        out[0] = H->comm.world;
        out[1] = H->comm.rank;
        out[2] = H->p2p.on ? 1 : 0;
        int nd = 0;
        for (auto& l : H->L) nd += l.sp.dist ? 1 : 0;
        out[3] = nd;
    });
    MGB_CATCH
}

int mgb200_adjust_nrhs(mgb200_handle h, int nrhs) {
    MGB_TRY
    MGB_BOTH(h, H->adjust_nrhs(nrhs));
    MGB_CATCH
}

int mgb200_set_cycle(mgb200_handle h, char cycle_type, const int64_t* relax_pre, const int64_t* relax_post) {
    MGB_TRY
    MGB_BOTH(h, H->set_cycle(cycle_type, relax_pre, relax_post));
    MGB_CATCH
}

int mgb200_cycle(mgb200_handle h, const void* b, void* x) {
    MGB_TRY
    MGB_CHECK(b && x, "null vector");
    MGB_BOTH(h, {
        H->ensure_work();
        const long long n = H->L[0].n;
        H->h2d_vec(b, H->L[0].b, n);
        H->h2d_vec(x, H->ucur, n);
        const bool xzero = (H->norm(n * H->m, H->ucur) == 0.0);  // `if norm(x)>0.0` (MGcycle.jl:29)
        H->cycle_top(xzero);
        H->d2h_vec(H->ucur, x, n);
    });
    MGB_CATCH
}

int mgb200_precondition(mgb200_handle h, const void* r, void* z) {
    MGB_TRY
    MGB_CHECK(r && z, "null vector");
    MGB_BOTH(h, {
        H->ensure_work();
        const long long n = H->L[0].n;
        H->h2d_vec(r, H->L[0].b, n);
        H->cycle_top(true);      // z .= 0 is a flag, not a copy: the first sweep from zero is x = d .* b
        H->d2h_vec(H->ucur, z, n);
    });
    MGB_CATCH
}

int mgb200_solveMG(mgb200_handle h, const void* b, void* x, double tol, int max_iter, int* iter, double* resvec) {
    MGB_TRY
    MGB_CHECK(b && x && iter && resvec, "null argument");
    MGB_BOTH(h, {
        H->ensure_work();
        const long long n = H->L[0].n;
        H->h2d_vec(b, H->L[0].b, n);
        H->h2d_vec(x, H->ucur, n);
        *iter = H->solveMG(tol, max_iter, resvec);
        H->d2h_vec(H->ucur, x, n);
    });
    MGB_CATCH
}

int mgb200_solveCG(mgb200_handle h, const void* b, void* x, double tol, int max_iter, int* iter, int* flag,
                   double* resvec) {
    MGB_TRY
    MGB_CHECK(b && x && iter && flag && resvec, "null argument");
    MGB_BOTH(h, {
        H->ensure_work();
        const long long n = H->L[0].n;
        H->h2d_vec(b, H->L[0].b, n);
        H->h2d_vec(x, H->ucur, n);
        if (H->m == 1) *iter = H->solveCG(H->ucur, tol, max_iter, flag, resvec);
        else *iter = H->solveBlockCG(H->ucur, tol, max_iter, flag, resvec, -1.0);
        H->d2h_vec(H->ucur, x, n);
    });
    MGB_CATCH
}

int mgb200_solveFGMRES(mgb200_handle h, const void* b, void* x, int inner, int flexible, double tol,
                       int max_iter, int* iter, int* flag, double* resvec, int* nres) {
    MGB_TRY
    MGB_CHECK(b && x && iter && flag && resvec && nres, "null argument");
    MGB_BOTH(h, {
        H->ensure_work();
        const long long n = H->L[0].n;
        H->h2d_vec(b, H->L[0].b, n);
        H->h2d_vec(x, H->ucur, n);
        *iter = H->solveFGMRES(H->ucur, inner, flexible != 0, tol, max_iter, flag, resvec, nres);
        H->d2h_vec(H->ucur, x, n);
    });
    MGB_CATCH
}

int mgb200_solveBiCGSTAB(mgb200_handle h, const void* b, void* x, double tol, int max_iter, int* iter, int* flag,
                         double* resvec, int* nprec) {
    MGB_TRY
    MGB_CHECK(b && x && iter && flag && resvec && nprec, "null argument");
    MGB_BOTH(h, {
        H->ensure_work();
        const long long n = H->L[0].n;
        H->h2d_vec(b, H->L[0].b, n);
        H->h2d_vec(x, H->ucur, n);
        *iter = H->solveBiCGSTAB(H->ucur, tol, max_iter, flag, resvec, nprec);
        H->d2h_vec(H->ucur, x, n);
    });
    MGB_CATCH
}

int mgb200_spmatmul(mgb200_handle h, int level, int which, double alpha, const void* x, double beta, void* y) {
    MGB_TRY
    MGB_CHECK(x && y, "null vector");
    MGB_BOTH(h, spmatmul_impl(H, level, which, alpha, x, beta, y));
    MGB_CATCH
}

int mgb200_device_buffers(mgb200_handle h, void** d_b, void** d_x) {
    MGB_TRY
    MGB_BOTH(h, {
        H->ensure_work();
        if (d_b) *d_b = H->L[0].b;
        if (d_x) *d_x = H->ucur;
    });
    MGB_CATCH
}

int mgb200_cycle_device(mgb200_handle h, int x_is_zero, void** d_x_out) {
    MGB_TRY
    MGB_BOTH(h, {
        void* p = H->cycle_top(x_is_zero != 0);
        if (d_x_out) *d_x_out = p;
    });
    MGB_CATCH
}

int mgb200_solveMG_device(mgb200_handle h, double tol, int max_iter, int* iter, double* resvec) {
    MGB_TRY
    MGB_CHECK(iter && resvec, "null argument");
    MGB_BOTH(h, *iter = H->solveMG(tol, max_iter, resvec));
    MGB_CATCH
}

int mgb200_solveCG_device(mgb200_handle h, double tol, int max_iter, int* iter, int* flag, double* resvec) {
    MGB_TRY
    MGB_CHECK(iter && flag && resvec, "null argument");
    MGB_BOTH(h, {
        H->ensure_work();
        if (H->m == 1) *iter = H->solveCG(H->ucur, tol, max_iter, flag, resvec);
        else *iter = H->solveBlockCG(H->ucur, tol, max_iter, flag, resvec, -1.0);
    });
    MGB_CATCH
}

int mgb200_synchronize(mgb200_handle h) {
    MGB_TRY
    MGB_BOTH(h, H->ctx.sync());
    MGB_CATCH
}

int mgb200_kernel_config(mgb200_handle h, int level, int which, int64_t* out) {
    MGB_TRY
    MGB_CHECK(out, "null output");
    MGB_BOTH(h, {
        MGB_CHECK(level >= 1 && level <= H->levels, "level out of range");
        auto& lv = H->L[level - 1];
        auto fill = [&](auto& M) {
            out[0] = M.tpr;
            out[1] = M.rpc;
            out[2] = (int64_t)M.smem;
            out[3] = M.staged ? 1 : 0;
            out[4] = M.nnz;
            out[5] = M.max_len;
        };
        if (which == 0) fill(lv.A);
        else if (which == 1) fill(lv.P);
        else fill(lv.R);
    });
    MGB_CATCH
}

int mgb200_pattern_info(mgb200_handle h, int level, int which, int64_t* out) {
    MGB_TRY
    MGB_CHECK(out, "null output");
    MGB_BOTH(h, {
        MGB_CHECK(level >= 1 && level <= H->levels, "level out of range");
        auto& lv = H->L[level - 1];
        auto fill = [&](auto& M) {
            out[0] = (M.pat.present && H->ctx.use_patterns) ? 1 : 0;
            out[1] = M.pat.rowrel ? 1 : 0;
            out[2] = M.pat.npat;
            out[3] = M.pat.nent;
        };
        if (which == 0) fill(lv.A);
        else if (which == 1) fill(lv.P);
        else fill(lv.R);
        out[4] = (which == 0 && lv.dpat != nullptr) ? 1 : 0;
    });
    MGB_CATCH
}

int mgb200_set_option(mgb200_handle h, const char* key, int64_t value) {
    MGB_TRY
    MGB_CHECK(key, "null key");
    MGB_BOTH(h, {
        const std::string k(key);
        if (k == "patterns") H->ctx.use_patterns = (int)value;
        else if (k == "graphs") H->ctx.use_graphs = (int)value;
        else if (k == "smem_budget") H->ctx.smem_budget = (int)value;
        else if (k == "tma") H->ctx.use_tma = (int)value;
        else if (k == "tma_min_rows") H->ctx.tma_min_rows = (int)value;
        else if (k == "box") H->ctx.use_box = (int)value;
        else if (k == "box_variant") H->ctx.box_variant = (int)value;
        else if (k == "box_variant27") H->ctx.box_variant27 = (int)value;
        else if (k == "box_variant_c") H->ctx.box_variant_c = (int)value;
        else if (k == "fuse_first") H->ctx.fuse_first_sweeps = (int)value;
        else if (k == "overlap_box") H->ctx.overlap_box = (int)value;
        else if (k == "box_min_rows") H->ctx.box_min_rows = (int)value;
        else if (k == "grid_transfers") H->ctx.grid_transfers = (int)value;
        else if (k == "gxp_quad") H->ctx.gxp_quad = (int)value;
        else if (k == "mrhs_march") H->ctx.mrhs_march = (int)value;
        else if (k == "split_test") H->ctx.split_test = (int)value;
        else if (k == "overlap") H->ctx.use_overlap = (int)value;
        else if (k == "fused_put") H->use_fused_put = (int)value;
        else throw Error(-1, "mgb200_set_option: unknown key " + k);
        H->invalidate_graphs();
    });
    MGB_CATCH
}

int mgb200_host_build_patterns(int64_t n_rows, const int64_t* colptr, const int64_t* rowval, const double* nzval,
                               int index_base, int max_patterns, int max_entries, int64_t* info, uint16_t* pid,
                               int32_t* c0, int32_t* pat_off, int32_t* delta, double* val) {
    MGB_TRY
    MGB_CHECK(colptr && rowval && nzval && info && pid && c0 && pat_off && delta && val, "null argument");
    HostPatterns<double> hp;
    const bool ok = build_patterns<double>(n_rows, colptr, rowval, nzval, index_base, false, max_patterns,
                                           max_entries, hp);
    info[0] = ok ? 1 : 0;
    info[1] = info[2] = info[3] = 0;
    if (ok) {
        info[1] = hp.rowrel ? 1 : 0;
        info[2] = hp.npat();
        info[3] = (int64_t)hp.delta.size();
        std::memcpy(pid, hp.pid.data(), n_rows * sizeof(uint16_t));
        if (!hp.rowrel) std::memcpy(c0, hp.c0.data(), n_rows * sizeof(int32_t));
        std::memcpy(pat_off, hp.pat_off.data(), hp.pat_off.size() * sizeof(int32_t));
        std::memcpy(delta, hp.delta.data(), hp.delta.size() * sizeof(int32_t));
        std::memcpy(val, hp.val.data(), hp.val.size() * sizeof(double));
    }
    MGB_CATCH
}

int mgb200_host_tma_plan(int64_t n_rows, const int64_t* colptr, const int64_t* rowval, const double* nzval,
                         int index_base, int tile, int elem_bytes, int max_patterns, int max_entries, int64_t* info,
                         int32_t* win, int32_t* delta, int32_t* soff) {
    MGB_TRY
    MGB_CHECK(colptr && rowval && nzval && info && win && delta && soff && tile > 0, "bad argument");
    MGB_CHECK(elem_bytes == 4 || elem_bytes == 8 || elem_bytes == 16, "elem_bytes must be 4, 8 or 16");
    HostPatterns<double> hp;
    const bool ok = build_patterns<double>(n_rows, colptr, rowval, nzval, index_base, false, max_patterns,
                                           max_entries, hp);
    info[0] = info[1] = info[2] = info[3] = 0;
    info[4] = ok ? (int64_t)hp.delta.size() : 0;
    if (ok && hp.rowrel) {
        if (elem_bytes == 4) host_tma_plan_impl<float>(hp, tile, info, win, delta, soff);
        else if (elem_bytes == 8) host_tma_plan_impl<double>(hp, tile, info, win, delta, soff);
        else host_tma_plan_impl<cplx>(hp, tile, info, win, delta, soff);
    }
    MGB_CATCH
}

int mgb200_host_detect_box(int64_t n_rows, const int64_t* colptr, const int64_t* rowval, const double* nzval,
                           int index_base, int max_patterns, int max_entries, int64_t* info, int32_t* mask) {
    MGB_TRY
    MGB_CHECK(colptr && rowval && nzval && info && mask, "null argument");
    HostPatterns<double> hp;
    const bool ok = build_patterns<double>(n_rows, colptr, rowval, nzval, index_base, false, max_patterns,
                                           max_entries, hp);
    info[0] = info[1] = info[2] = info[3] = 0;
    if (ok) {
        const BoxInfo B = detect_box<double>(hp, n_rows);
        if (B.ok) {
            info[0] = 1;
            info[1] = B.S;
            info[2] = B.S2;
            info[3] = hp.npat();
            for (int p = 0; p < hp.npat(); ++p) mask[p] = B.mask[p];
        }
    }
    MGB_CATCH
}

int mgb200_host_box_apply(int64_t n_rows, const int64_t* colptr, const int64_t* rowval, const double* nzval,
                          int index_base, int mode, int rows_per_thread, int base_rows, int ctas, int fold_d, const double* x,
                          const double* b, const double* d, double* y, int64_t* info) {
    MGB_TRY
    MGB_CHECK(colptr && rowval && nzval && x && y && info, "null argument");
    MGB_CHECK(mode == 0 || mode == 2 || mode == 3 || mode == 4, "mode must be 0, 2, 3 or 4");
    MGB_CHECK(rows_per_thread == 1 || rows_per_thread == 2 || rows_per_thread == 4, "rows_per_thread must be 1, 2 or 4");
    MGB_CHECK(base_rows >= 8 && base_rows % 8 == 0 && ctas >= 1, "base_rows must be a multiple of 8");
    MGB_CHECK(mode == 0 || b, "b required");
    MGB_CHECK(mode != 3 || d, "d required");
    info[0] = info[1] = info[2] = info[3] = 0;
    HostPatterns<double> hp;
    if (!build_patterns<double>(n_rows, colptr, rowval, nzval, index_base, false, PAT_MAX_PATTERNS, PAT_MAX_ENTRIES, hp)) return 0;
    if (!hp.rowrel) return 0;
    const BoxInfo B = detect_box<double>(hp, n_rows);
    int shape = 0, NP = 0, p0 = 0;
    std::vector<double> ctab;
    BoxCoef<double> C0;
    if (!box_build_tables<double>(hp, B, n_rows, shape, NP, p0, ctab, C0)) return 0;
    const int RZ = rows_per_thread;
    if (RZ > 1 && n_rows % B.S2 != 0) return 0;
    std::vector<double> dtab(NP, 0.0);
    bool dp = false;
    if ((mode == 3 && fold_d) || mode == 4) {
        dp = true;
        for (int p = 0; p < hp.npat(); ++p) dtab[p] = d[hp.rep_row[p]];
        C0.d0 = dtab[p0];
    }
    BoxPlan P;
    box_make_plan<double>(P, shape, RZ, base_rows, n_rows, B.S, B.S2, 0, (n_rows + 1) & ~1LL, hp.npat(), p0);
    long long fast_rows = 0;
    std::vector<uint16_t> pidp(hp.pid.size() + 16, 0);       // the device array carries 16 zeroed ids of slack
    std::copy(hp.pid.begin(), hp.pid.end(), pidp.begin());
#define MGB_HB(SHAPE, RR)                                                                                                   \
    {                                                                                                                       \
        if (mode == 0) host_box_run<SHAPE, 0, false, RR>(P, C0, ctab, dtab, base_rows, ctas, pidp.data(), x, b, d, y, &fast_rows);      \
        else if (mode == 2) host_box_run<SHAPE, 2, false, RR>(P, C0, ctab, dtab, base_rows, ctas, pidp.data(), x, b, d, y, &fast_rows); \
        else if (mode == 4) host_box_run<SHAPE, 4, true, RR>(P, C0, ctab, dtab, base_rows, ctas, pidp.data(), x, b, d, y, &fast_rows);  \
        else if (dp) host_box_run<SHAPE, 3, true, RR>(P, C0, ctab, dtab, base_rows, ctas, pidp.data(), x, b, d, y, &fast_rows);         \
        else host_box_run<SHAPE, 3, false, RR>(P, C0, ctab, dtab, base_rows, ctas, pidp.data(), x, b, d, y, &fast_rows);                \
    }
#define MGB_HBR(SHAPE) { if (RZ == 1) MGB_HB(SHAPE, 1) else if (RZ == 2) MGB_HB(SHAPE, 2) else MGB_HB(SHAPE, 4) }
    if (shape == 7) MGB_HBR(7) else MGB_HBR(27)
#undef MGB_HBR
#undef MGB_HB
    info[0] = 1;
    info[1] = shape;
    info[2] = hp.npat();
    info[3] = fast_rows;
    MGB_CATCH
}

int mgb200_host_pattern_apply(int64_t n_rows, const int64_t* colptr, const int64_t* rowval, const double* nzval,
                              int index_base, int mode, int fold_d, const double* x, const double* b, const double* d,
                              double* y, int64_t* info) {
    MGB_TRY
    MGB_CHECK(colptr && rowval && nzval && x && y && info, "null argument");
    MGB_CHECK(mode == 0 || mode == 2 || mode == 3, "mode must be 0, 2 or 3");
    MGB_CHECK(mode == 0 || b, "b required");
    MGB_CHECK(mode != 3 || d, "d required");
    info[0] = info[1] = info[2] = info[3] = 0;
    HostPatterns<double> hp;
    if (!build_patterns<double>(n_rows, colptr, rowval, nzval, index_base, false, PAT_MAX_PATTERNS, PAT_MAX_ENTRIES, hp)) return 0;
    if (!hp.rowrel) return 0;
    const BoxInfo B = detect_box<double>(hp, n_rows);      // S2 only steers which of the three x pointers a walk uses
    std::vector<PatEntry<double>> ent(hp.delta.size());
    for (size_t k = 0; k < ent.size(); ++k) {
        ent[k].v = hp.val[k];
        ent[k].delta = hp.delta[k];
    }
    std::vector<double> dpat(hp.npat(), 0.0);
    const bool dp = (mode == 3 && fold_d);
    if (dp)
        for (int p = 0; p < hp.npat(); ++p) dpat[p] = d[hp.rep_row[p]];
    const long long S2 = B.ok ? B.S2 : 0;
    if (mode == 0) host_pattern_walk<0, false>(hp, S2, n_rows, ent, nullptr, x, b, d, y);
    else if (mode == 2) host_pattern_walk<2, false>(hp, S2, n_rows, ent, nullptr, x, b, d, y);
    else if (dp) host_pattern_walk<3, true>(hp, S2, n_rows, ent, dpat.data(), x, b, d, y);
    else host_pattern_walk<3, false>(hp, S2, n_rows, ent, nullptr, x, b, d, y);
    info[0] = 1;
    info[1] = B.ok ? B.S : 0;
    info[2] = S2;
    MGB_CATCH
}

int mgb200_host_grid_transfer(int kind, int dim, const int64_t* n_fine_nodes, const int64_t* n_coarse_nodes, int64_t n_rows,
                              const int64_t* colptr, const int64_t* rowval, const double* nzval, int index_base,
                              int lines_per_thread, const double* x, double* y, int64_t* info) {
    MGB_TRY
    MGB_CHECK(colptr && rowval && nzval && x && y && info && n_fine_nodes && n_coarse_nodes, "null argument");
    MGB_CHECK(kind == 1 || kind == 2, "kind must be 1 (prolongation) or 2 (restriction)");
    MGB_CHECK(dim >= 1 && dim <= 3, "dim must be 1, 2 or 3");
    MGB_CHECK(lines_per_thread == 0 || lines_per_thread == 1 || lines_per_thread == 2 || lines_per_thread == 4,
              "lines_per_thread must be 0, 1, 2 or 4");
    info[0] = 0;
    HostPatterns<double> hp;
    if (!build_patterns<double>(n_rows, colptr, rowval, nzval, index_base, false, PAT_MAX_PATTERNS, PAT_MAX_ENTRIES, hp)) return 0;
    int n[3], N[3];
    for (int d = 0; d < 3; ++d) {
        n[d] = d < dim ? (int)n_fine_nodes[d] : 1;
        N[d] = d < dim ? (int)n_coarse_nodes[d] : 1;
    }
    GridXfer X;
    const bool ok = kind == 1 ? gx_verify_prolongation<double>(hp, n_rows, n, N, X) : gx_verify_restriction<double>(hp, n_rows, n, N, X);
    if (!ok) return 0;
    std::vector<PatEntry<double>> ent(hp.delta.size());
    for (size_t k = 0; k < ent.size(); ++k) {
        ent[k].v = hp.val[k];
        ent[k].delta = hp.delta[k];
    }
    if (lines_per_thread == 0) {           // dictionary walk, one row at a time: column = c0[row] + delta
        for (long long row = 0; row < n_rows; ++row) {
            const int p = hp.pid[row];
            double acc = 0.0;
            for (int k = hp.pat_off[p]; k < hp.pat_off[p + 1]; ++k) acc = acc + hp.val[k] * x[hp.c0[row] + hp.delta[k]];
            y[row] = kind == 1 ? y[row] + acc : acc;
        }
    } else {
        host_gx_run(X, gx_dense_table<double>(X, hp.val), x, y);
    }
    info[0] = 1;
    MGB_CATCH
}

int mgb200_profile_enable(mgb200_handle h, int on) {
    MGB_TRY
    MGB_BOTH(h, {
        H->ctx.sync();
        H->ctx.profiling = (on != 0);
    });
    MGB_CATCH
}

int mgb200_profile_report(mgb200_handle h, double* records, int max_records, int* nrec) {
    MGB_TRY
    MGB_CHECK(records && nrec, "null argument");
    MGB_BOTH(h, {
        Context& c = H->ctx;
        c.sync();
        std::map<std::pair<int, int>, std::array<double, 4>> agg;
        for (auto& r : c.prof) {
            float ms = 0.f;
            MGB_CUDA(cudaEventElapsedTime(&ms, r.e0, r.e1));
            auto& a = agg[std::make_pair(r.kind, r.level)];
            a[0] += 1.0;
            a[1] += ms;
            a[2] += r.bytes;
            a[3] += r.fmt_bytes;
            c.ev_pool.push_back(r.e0);
            c.ev_pool.push_back(r.e1);
        }
        c.prof.clear();
        int k = 0;
        for (auto& kv : agg) {
            if (k >= max_records) break;
            records[6 * k + 0] = kv.first.first;
            records[6 * k + 1] = kv.first.second;
            records[6 * k + 2] = kv.second[0];
            records[6 * k + 3] = kv.second[1];
            records[6 * k + 4] = kv.second[2];
            records[6 * k + 5] = kv.second[3];
            ++k;
        }
        *nrec = k;
    });
    MGB_CATCH
}

int mgb200_event_record(mgb200_handle h, int idx) {
    MGB_TRY
    MGB_CHECK(idx >= 0 && idx < 16, "event slot out of range");
    MGB_BOTH(h, {
        Context& c = H->ctx;
        if (c.user_ev.empty()) {
            c.user_ev.resize(16);
            for (auto& e : c.user_ev) MGB_CUDA(cudaEventCreate(&e));
        }
        MGB_CUDA(cudaEventRecord(c.user_ev[idx], c.stream));
    });
    MGB_CATCH
}

int mgb200_event_elapsed_ms(mgb200_handle h, int i0, int i1, double* ms) {
    MGB_TRY
    MGB_CHECK(ms && i0 >= 0 && i0 < 16 && i1 >= 0 && i1 < 16, "bad event arguments");
    MGB_BOTH(h, {
        Context& c = H->ctx;
        MGB_CHECK(!c.user_ev.empty(), "no event recorded");
        MGB_CUDA(cudaEventSynchronize(c.user_ev[i1]));
        float f = 0.f;
        MGB_CUDA(cudaEventElapsedTime(&f, c.user_ev[i0], c.user_ev[i1]));
        *ms = f;
    });
    MGB_CATCH
}

// Stand-alone y = A_level x on device-resident buffers, timed with CUDA events on the library's stream
// (BASELINE metric "SpMV GB/s vs HBM peak"): x = memCycle[level].b, y = memCycle[level].r.
int mgb200_bench_spmv(mgb200_handle h, int level, int reps, double* ms_per_call, double* algorithmic_bytes) {
    MGB_TRY
    MGB_CHECK(ms_per_call && algorithmic_bytes && reps >= 1, "bad argument");
    MGB_BOTH(h, {
        H->ensure_work();
        MGB_CHECK(level >= 1 && level < H->levels, "level out of range");
        auto& lv = H->L[level - 1];
        Context& c = H->ctx;
        typedef typename std::remove_reference<decltype(*lv.b)>::type TV;
        for (int i = 0; i < 3; ++i) H->apply_A(lv.A, lv.b, lv.r, level);
        cudaEvent_t e0 = c.get_event(), e1 = c.get_event();
        MGB_CUDA(cudaEventRecord(e0, c.stream));
        for (int i = 0; i < reps; ++i) H->apply_A(lv.A, lv.b, lv.r, level);
        MGB_CUDA(cudaEventRecord(e1, c.stream));
        MGB_CUDA(cudaEventSynchronize(e1));
        float f = 0.f;
        MGB_CUDA(cudaEventElapsedTime(&f, e0, e1));
        c.ev_pool.push_back(e0);
        c.ev_pool.push_back(e1);
        *ms_per_call = f / reps;
        *algorithmic_bytes = csr_bytes<TV, TV>(lv.A, MODE_SPMV, H->m);
    });
    MGB_CATCH
}

int mgb200_profiler_start(void) {
    MGB_TRY
    MGB_CUDA(cudaProfilerStart());
    MGB_CATCH
}
int mgb200_profiler_stop(void) {
    MGB_TRY
    MGB_CUDA(cudaProfilerStop());
    MGB_CATCH
}

int64_t mgb200_launch_count(mgb200_handle h) {
    if (!h || !h->impl) return -1;
    return h->impl->ctx.launches;
}

// ---- host-only helpers exported for the CPU test-suite (no GPU needed) ---------------------------
// t = pinv(H) xi for a Hermitian n x n H (row-major, interleaved re/im)
int mgb200_host_pinv_apply(int n, const double* H, const double* xi, double* t) {
    MGB_TRY
    std::vector<zc> Hv((size_t)n * n), xv(n), tv;
    for (int i = 0; i < n * n; ++i) Hv[i] = zc(H[2 * i], H[2 * i + 1]);
    for (int i = 0; i < n; ++i) xv[i] = zc(xi[2 * i], xi[2 * i + 1]);
    hermitian_pinv_apply(n, Hv, xv, tv);
    for (int i = 0; i < n; ++i) {
        t[2 * i] = tv[i].real();
        t[2 * i + 1] = tv[i].imag();
    }
    MGB_CATCH
}
// P = pinv(A) for a general n x n complex matrix (row-major, interleaved re/im), cut-off rtol*max(sigma)
int mgb200_host_general_pinv(int n, const double* A, double rtol, double* P) {
    MGB_TRY
    std::vector<zc> Av((size_t)n * n), Pv;
    for (int i = 0; i < n * n; ++i) Av[i] = zc(A[2 * i], A[2 * i + 1]);
    general_pinv(n, Av, rtol, Pv);
    for (int i = 0; i < n * n; ++i) {
        P[2 * i] = Pv[i].real();
        P[2 * i + 1] = Pv[i].imag();
    }
    MGB_CATCH
}
// least squares on a (cols+1) x cols Hessenberg block (row-major, ld = cols); returns the residual in *res
int mgb200_host_hessenberg_lsq(int cols, const double* H, const double* xi, double* y, double* res) {
    MGB_TRY
    std::vector<zc> Hv((size_t)(cols + 1) * cols), xv(cols + 1), yv;
    for (int i = 0; i < (cols + 1) * cols; ++i) Hv[i] = zc(H[2 * i], H[2 * i + 1]);
    for (int i = 0; i < cols + 1; ++i) xv[i] = zc(xi[2 * i], xi[2 * i + 1]);
    *res = hessenberg_lsq(Hv, cols, cols + 1, cols, xv, yv);
    for (int i = 0; i < cols; ++i) {
        y[2 * i] = yv[i].real();
        y[2 * i + 1] = yv[i].imag();
    }
    MGB_CATCH
}

// block-Hessenberg least squares min || A Y - B ||_F (rows x cols, nb right-hand sides); *res = residual norm
int mgb200_host_dense_lsq(int rows, int cols, int nb, const double* A, const double* B, double* Y, double* res) {
    MGB_TRY
    MGB_CHECK(A && B && Y && res && rows >= cols && cols >= 1 && nb >= 1, "bad argument");
    std::vector<zc> Av((size_t)rows * cols), Bv((size_t)rows * nb), Yv;
    for (size_t i = 0; i < Av.size(); ++i) Av[i] = zc(A[2 * i], A[2 * i + 1]);
    for (size_t i = 0; i < Bv.size(); ++i) Bv[i] = zc(B[2 * i], B[2 * i + 1]);
    *res = dense_lsq(rows, cols, nb, Av, Bv, Yv);
    for (size_t i = 0; i < Yv.size(); ++i) {
        Y[2 * i] = Yv[i].real();
        Y[2 * i + 1] = Yv[i].imag();
    }
    MGB_CATCH
}
// what = 0: R = upper Cholesky factor of the Hermitian positive definite A (out: m x m); returns status -5 if A is not
// positive definite.  what = 1: out = A^{-1} B by LU with partial pivoting (B, out: m x nb); -5 if singular.
int mgb200_host_small_factor(int what, int m, int nb, const double* A, const double* B, double* out) {
    MGB_TRY
    MGB_CHECK(A && out && m >= 1, "bad argument");
    std::vector<zc> Av((size_t)m * m), Ov;
    for (size_t i = 0; i < Av.size(); ++i) Av[i] = zc(A[2 * i], A[2 * i + 1]);
    if (what == 0) {
        if (!cholesky_upper(m, Av, Ov)) throw Error(-5, "matrix is not positive definite");
    } else {
        MGB_CHECK(B && nb >= 1, "bad argument");
        std::vector<zc> Bv((size_t)m * nb);
        for (size_t i = 0; i < Bv.size(); ++i) Bv[i] = zc(B[2 * i], B[2 * i + 1]);
        if (!lu_solve_small(m, nb, Av, Bv, Ov)) throw Error(-5, "matrix is singular");
    }
    for (size_t i = 0; i < Ov.size(); ++i) {
        out[2 * i] = Ov[i].real();
        out[2 * i + 1] = Ov[i].imag();
    }
    MGB_CATCH
}
// rows [out[0], out[1]) of a CSR slab (columns relative to the first owned row of the input vector: < 0 lower
// ghost, >= n_in_owned upper ghost) that read owned rows only - the rows that run beside the halo exchange
int mgb200_host_interior_rows(int64_t n_rows, const int64_t* rowptr, const int64_t* cols, int64_t n_in_owned,
                              int64_t* out) {
    MGB_TRY
    MGB_CHECK(rowptr && cols && out && n_rows >= 0, "bad argument");
    HostRows<double> H;
    H.n_rows = n_rows;
    H.rowptr.assign(rowptr, rowptr + n_rows + 1);
    H.col.assign(cols, cols + rowptr[n_rows]);
    int lo = 0, hi = 0;
    interior_rows(H, n_in_owned, lo, hi);
    out[0] = lo;
    out[1] = hi;
    MGB_CATCH
}


/* ---- single-process multi-GPU entry (multi.cuh) ---------------------------------------------------------------------- */
static size_t mgb_val_bytes(int val_type, bool real_part) {
    switch (val_type) {
        case MGB200_FP64: return 8;
        case MGB200_CFP64: return real_part ? 8 : 16;
        case MGB200_FP32: return 4;
        default: return real_part ? 4 : 8;
    }
}

int mgb200_multi_create(mgb200_multi_handle* out, int n_devices, const int* devices, int val_type, int levels, int nrhs,
                        char cycle_type, int relax_kind, const int64_t* relax_pre, const int64_t* relax_post) {
    MGB_TRY
    MGB_CHECK(out && n_devices >= 1 && n_devices <= 64, "bad device count");
    int have = 0;
    MGB_CUDA(cudaGetDeviceCount(&have));
    std::unique_ptr<MultiHandle> M(new MultiHandle());
    M->G = n_devices;
    M->val_type = val_type;
    M->levels = levels;
    M->nrhs = nrhs;
    for (int g = 0; g < n_devices; ++g) {
        const int dev = devices ? devices[g] : g;
        MGB_CHECK(dev >= 0 && dev < have, "device index out of range");
        M->devices.push_back(dev);
    }
    M->h.assign(n_devices, nullptr);
    char id[128];
    if (n_devices > 1) {
        const int st = mgb200_dist_unique_id(id);
        if (st != 0) return st;
    }
    MultiHandle* Mp = M.get();
    const int st = multi_run(Mp, [&](int g) {
        int s = mgb200_create(&Mp->h[g], val_type, levels, nrhs, cycle_type, relax_kind, relax_pre, relax_post, Mp->devices[g]);
        if (s == 0) s = mgb200_dist_init(Mp->h[g], g, Mp->G, Mp->G > 1 ? id : nullptr);
        return s;
    });
    if (st != 0) {
        g_last_error = Mp->error;
        for (auto h : Mp->h)
            if (h) mgb200_destroy(h);
        return st;
    }
    *out = reinterpret_cast<mgb200_multi_handle>(M.release());
    MGB_CATCH
}

int mgb200_multi_destroy(mgb200_multi_handle mh) {
    MGB_TRY
    MultiHandle* M = reinterpret_cast<MultiHandle*>(mh);
    if (!M) return 0;
    multi_run(M, [&](int g) { return M->h[g] ? mgb200_destroy(M->h[g]) : 0; });
    delete M;
    MGB_CATCH
}

int mgb200_multi_upload_level(mgb200_multi_handle mh, int level, int64_t n, int64_t nc, const int64_t* row_offsets,
                              const int64_t* coarse_row_offsets, const int64_t* a_colptr, const int64_t* a_rowval,
                              const void* a_nzval, const int64_t* p_colptr, const int64_t* p_rowval, const void* p_nzval,
                              const int64_t* r_colptr, const int64_t* r_rowval, const void* r_nzval, const void* d, int index_base) {
    MGB_TRY
    MultiHandle* M = reinterpret_cast<MultiHandle*>(mh);
    MGB_CHECK(M, "null handle");
    MGB_CHECK(a_colptr && a_rowval && a_nzval && p_colptr && p_rowval && p_nzval && r_colptr && r_rowval && r_nzval && d, "null array");
    int st;
    if (!row_offsets || M->G == 1) {
        // replicated level (or one device): every rank holds the global matrices
        st = multi_run(M, [&](int g) {
            return mgb200_upload_level(M->h[g], level, n, nc, a_colptr, a_rowval, a_nzval, p_colptr, p_rowval, p_nzval, r_colptr,
                                       r_rowval, r_nzval, d, index_base);
        });
        if (level == 1) {
            M->fine_offsets.assign(M->G + 1, 0);
            M->fine_offsets[M->G] = n;          // every rank works on the whole vector (G == 1, or nothing is partitioned)
        }
    } else {
        MGB_CHECK(coarse_row_offsets, "coarse_row_offsets required for a row-partitioned level");
        MGB_CHECK(row_offsets[0] == 0 && row_offsets[M->G] == n && coarse_row_offsets[0] == 0 && coarse_row_offsets[M->G] == nc,
                  "row offsets must cover [0, n] and [0, nc]");
        const size_t va = mgb_val_bytes(M->val_type, false), vr = mgb_val_bytes(M->val_type, true);
        if (level == 1) M->fine_offsets.assign(row_offsets, row_offsets + M->G + 1);
        st = multi_run(M, [&](int g) {
            const int64_t lo = row_offsets[g], hi = row_offsets[g + 1], clo = coarse_row_offsets[g], chi = coarse_row_offsets[g + 1];
            const CscSlice A = csc_columns(a_colptr, a_rowval, a_nzval, va, lo, hi, index_base);
            const CscSlice P = csc_columns(p_colptr, p_rowval, p_nzval, vr, lo, hi, index_base);
            const CscSlice R = csc_columns(r_colptr, r_rowval, r_nzval, vr, clo, chi, index_base);
            return mgb200_dist_upload_level(M->h[g], level, n, row_offsets, nc, coarse_row_offsets, A.colptr.data(), A.rowval, A.nzval,
                                            P.colptr.data(), P.rowval, P.nzval, R.colptr.data(), R.rowval, R.nzval,
                                            static_cast<const unsigned char*>(d) + (size_t)lo * va, index_base);
        });
    }
    if (st != 0) g_last_error = M->error;
    return st;
    MGB_CATCH
}

int mgb200_multi_set_level_grid(mgb200_multi_handle mh, int level, int dim, const int64_t* n_fine_nodes, const int64_t* n_coarse_nodes) {
    MGB_TRY
    MultiHandle* M = reinterpret_cast<MultiHandle*>(mh);
    MGB_CHECK(M, "null handle");
    for (int g = 0; g < M->G; ++g) {
        const int st = mgb200_set_level_grid(M->h[g], level, dim, n_fine_nodes, n_coarse_nodes);
        if (st != 0) return st;
    }
    MGB_CATCH
}

int mgb200_multi_upload_coarsest(mgb200_multi_handle mh, int64_t n, const int64_t* colptr, const int64_t* rowval, const void* nzval,
                                 int index_base) {
    MGB_TRY
    MultiHandle* M = reinterpret_cast<MultiHandle*>(mh);
    MGB_CHECK(M, "null handle");
    const int st = multi_run(M, [&](int g) { return mgb200_upload_coarsest(M->h[g], n, colptr, rowval, nzval, index_base); });
    if (st != 0) g_last_error = M->error;
    return st;
    MGB_CATCH
}

/* the solve entry points: every rank gets its rows of b and x (nrhs = 1: contiguous slices of the global vectors) */
static int multi_solve(mgb200_multi_handle mh, const void* b, void* x, const std::function<int(int, const void*, void*)>& call) {
    MultiHandle* M = reinterpret_cast<MultiHandle*>(mh);
    if (!M) {
        g_last_error = "mgb200: null handle";
        return -1;
    }
    if (M->fine_offsets.size() != (size_t)M->G + 1 || !b || !x) {
        g_last_error = "mgb200: level 1 not uploaded, or null vector";
        return -1;
    }
    if (M->nrhs != 1 && M->G > 1) {
        g_last_error = "mgb200_multi_*: the row-partitioned path takes one right-hand side";
        return -1;
    }
    const size_t va = mgb_val_bytes(M->val_type, false);
    const int st = multi_run(M, [&](int g) {
        const int64_t lo = (M->fine_offsets[g + 1] == M->fine_offsets[M->G] && M->fine_offsets[g] == 0 && g > 0) ? 0 : M->fine_offsets[g];
        return call(g, static_cast<const unsigned char*>(b) + (size_t)lo * va, static_cast<unsigned char*>(x) + (size_t)lo * va);
    });
    if (st != 0) g_last_error = M->error;
    return st;
}

int mgb200_multi_solveMG(mgb200_multi_handle mh, const void* b, void* x, double tol, int max_iter, int* iter, double* resvec) {
    MGB_TRY
    MGB_CHECK(iter && resvec, "null argument");
    MultiHandle* M = reinterpret_cast<MultiHandle*>(mh);
    MGB_CHECK(M, "null handle");
    std::vector<int> its(M->G, 0);
    std::vector<std::vector<double>> rv(M->G, std::vector<double>(max_iter + 1, 0.0));
    const int st = multi_solve(mh, b, x, [&](int g, const void* bg, void* xg) {
        return mgb200_solveMG(M->h[g], bg, xg, tol, max_iter, &its[g], rv[g].data());
    });
    if (st != 0) return st;
    *iter = its[0];
    std::copy(rv[0].begin(), rv[0].end(), resvec);
    MGB_CATCH
}

int mgb200_multi_solveCG(mgb200_multi_handle mh, const void* b, void* x, double tol, int max_iter, int* iter, int* flag,
                         double* resvec) {
    MGB_TRY
    MGB_CHECK(iter && flag && resvec, "null argument");
    MultiHandle* M = reinterpret_cast<MultiHandle*>(mh);
    MGB_CHECK(M, "null handle");
    std::vector<int> its(M->G, 0), fl(M->G, 0);
    std::vector<std::vector<double>> rv(M->G, std::vector<double>(std::max(max_iter, 1), 0.0));
    const int st = multi_solve(mh, b, x, [&](int g, const void* bg, void* xg) {
        return mgb200_solveCG(M->h[g], bg, xg, tol, max_iter, &its[g], &fl[g], rv[g].data());
    });
    if (st != 0) return st;
    *iter = its[0];
    *flag = fl[0];
    std::copy(rv[0].begin(), rv[0].end(), resvec);
    MGB_CATCH
}

int mgb200_multi_solveFGMRES(mgb200_multi_handle mh, const void* b, void* x, int inner, int flexible, double tol, int max_iter,
                             int* iter, int* flag, double* resvec, int* nres) {
    MGB_TRY
    MGB_CHECK(iter && flag && resvec && nres, "null argument");
    MultiHandle* M = reinterpret_cast<MultiHandle*>(mh);
    MGB_CHECK(M, "null handle");
    std::vector<int> its(M->G, 0), fl(M->G, 0), nr(M->G, 0);
    std::vector<std::vector<double>> rv(M->G, std::vector<double>(std::max(inner * max_iter, 1), 0.0));
    const int st = multi_solve(mh, b, x, [&](int g, const void* bg, void* xg) {
        return mgb200_solveFGMRES(M->h[g], bg, xg, inner, flexible, tol, max_iter, &its[g], &fl[g], rv[g].data(), &nr[g]);
    });
    if (st != 0) return st;
    *iter = its[0];
    *flag = fl[0];
    *nres = nr[0];
    std::copy(rv[0].begin(), rv[0].begin() + nr[0], resvec);
    MGB_CATCH
}

int mgb200_multi_precondition(mgb200_multi_handle mh, const void* r, void* z) {
    MGB_TRY
    MultiHandle* M = reinterpret_cast<MultiHandle*>(mh);
    MGB_CHECK(M, "null handle");
    return multi_solve(mh, r, z, [&](int g, const void* rg, void* zg) { return mgb200_precondition(M->h[g], rg, zg); });
    MGB_CATCH
}

int mgb200_multi_info(mgb200_multi_handle mh, int64_t* out) {
    MGB_TRY
    MultiHandle* M = reinterpret_cast<MultiHandle*>(mh);
    MGB_CHECK(M && out, "null argument");
    std::vector<std::vector<int64_t>> o(M->G, std::vector<int64_t>(4, 0));
    const int st = multi_run(M, [&](int g) { return mgb200_dist_info(M->h[g], o[g].data()); });
    if (st != 0) {
        g_last_error = M->error;
        return st;
    }
    for (int k = 0; k < 4; ++k) out[k] = o[0][k];
    MGB_CATCH
}

}  // extern "C"
