// Launch wrappers: pick the kernel instantiation chosen at upload and account the
// algorithmic bytes of SURVEY.md section 8(d) for every launch.
#pragma once
#include <type_traits>

#include "hierarchy.cuh"

namespace mgb200 {

// algorithmic bytes of one CSR pass (device format of the accounting: 4-byte column indices
// and row pointers, every vector read or written exactly once)
template <typename TA, typename TV>
static double csr_bytes(const Csr<TA>& M, int mode, int m) {
    double b = (double)M.nnz * (sizeof(TA) + 4) + 4.0 * (M.n_rows + 1);
    const double sv = sizeof(TV);
    switch (mode) {
        case MODE_SPMV: b += ((double)M.n_cols + M.n_rows) * sv * m; break;           // read x, write y
        case MODE_ADD: b += ((double)M.n_cols + 2.0 * M.n_rows) * sv * m; break;      // read x, read+write y
        case MODE_RESID: b += (3.0 * M.n_rows) * sv * m; break;                       // x, b, r
        case MODE_SWEEP: b += (3.0 * m + 1.0) * M.n_rows * sv; break;                 // x, b, x', d
    }
    return b;
}

template <typename TA, typename TV, int TPR>
static void launch_stream_mode(Context& ctx, const Csr<TA>& M, int mode, const TV* x, const TV* b, const TV* d,
                               TV* y) {
    const int grid = cdiv(M.n_rows, M.rpc);
    const int nt = M.rpc * TPR;
#define MGB_CASE(MODE)                                                                                   \
    case MODE: {                                                                                         \
        auto kern = csr_stream_kernel<TA, TV, TPR, MODE>;                                                \
        if (M.smem > 48 * 1024)                                                                          \
            MGB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx.max_smem_optin)); \
        kern<<<grid, nt, M.smem, ctx.stream>>>(M.n_rows, M.rowptr, M.colind, M.val, x, b, d, y, M.rpc,   \
                                               M.cap);                                                   \
    } break;
    switch (mode) {
        MGB_CASE(MODE_SPMV)
        MGB_CASE(MODE_ADD)
        MGB_CASE(MODE_RESID)
        MGB_CASE(MODE_SWEEP)
    }
#undef MGB_CASE
    MGB_LAUNCH_CHECK();
}

template <typename TA, typename TV>
static void launch_mrhs_mode(Context& ctx, const Csr<TA>& M, int mode, const TV* x, const TV* b, const TV* d,
                             TV* y, int m) {
    const int grid = cdiv(M.n_rows, M.rpc);
    int mp = 1;
    while (mp < m && mp < 32) mp *= 2;
    const int nt = 256;
#define MGB_CASE(MODE)                                                                                   \
    case MODE: {                                                                                         \
        auto kern = csr_stream_mrhs_kernel<TA, TV, MODE>;                                                \
        if (M.smem > 48 * 1024)                                                                          \
            MGB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx.max_smem_optin)); \
        kern<<<grid, nt, M.smem, ctx.stream>>>(M.n_rows, M.rowptr, M.colind, M.val, x, b, d, y, M.rpc,   \
                                               M.cap, m, mp);                                            \
    } break;
    switch (mode) {
        MGB_CASE(MODE_SPMV)
        MGB_CASE(MODE_ADD)
        MGB_CASE(MODE_RESID)
        MGB_CASE(MODE_SWEEP)
    }
#undef MGB_CASE
    MGB_LAUNCH_CHECK();
}

template <typename TA, typename TV>
static void launch_rowwarp_mode(Context& ctx, const Csr<TA>& M, int mode, const TV* x, const TV* b,
                                const TV* d, TV* y, int m) {
    const int grid = cdiv((long long)M.n_rows * 32, 256);
    switch (mode) {
        case MODE_SPMV:
            csr_rowwarp_kernel<TA, TV, MODE_SPMV><<<grid, 256, 0, ctx.stream>>>(M.n_rows, M.rowptr, M.colind, M.val, x, b, d, y, m);
            break;
        case MODE_ADD:
            csr_rowwarp_kernel<TA, TV, MODE_ADD><<<grid, 256, 0, ctx.stream>>>(M.n_rows, M.rowptr, M.colind, M.val, x, b, d, y, m);
            break;
        case MODE_RESID:
            csr_rowwarp_kernel<TA, TV, MODE_RESID><<<grid, 256, 0, ctx.stream>>>(M.n_rows, M.rowptr, M.colind, M.val, x, b, d, y, m);
            break;
        case MODE_SWEEP:
            csr_rowwarp_kernel<TA, TV, MODE_SWEEP><<<grid, 256, 0, ctx.stream>>>(M.n_rows, M.rowptr, M.colind, M.val, x, b, d, y, m);
            break;
    }
    MGB_LAUNCH_CHECK();
}

// vector bytes of one pass (the part of csr_bytes that is not the matrix)
template <typename TA, typename TV>
static double vec_bytes(const Csr<TA>& M, int mode, int m, bool d_from_dict) {
    const double sv = sizeof(TV);
    switch (mode) {
        case MODE_SPMV: return ((double)M.n_cols + M.n_rows) * sv * m;
        case MODE_ADD: return ((double)M.n_cols + 2.0 * M.n_rows) * sv * m;
        case MODE_RESID: return (3.0 * M.n_rows) * sv * m;
        default: return (3.0 * m + (d_from_dict ? 0.0 : 1.0)) * M.n_rows * sv;
    }
}

// stencil-dictionary kernel (pattern.cuh), one right-hand side
// TMA-staged variant (pat_tma_kernel): row-relative matrices, SPMV / RESID / SWEEP, large levels
// Tiles [t0, t1) of TmaTile::NT rows (t1 < 0: all).  CTAs worth `reserve_threads` threads fewer than the machine holds are launched
// when another kernel (the halo exchange of the same input vector, on its own stream) must find room beside the
// persistent CTAs of this one.
template <typename TA, typename TV>
static bool launch_pattern_tma(Context& ctx, const Csr<TA>& M, int mode, const TV* x, const TV* b, const TV* d,
                               const TV* dpat, TV* y, int t0 = 0, int t1 = -1, int reserve_threads = 0,
                               const PutPlan& pp = no_put()) {
    const PatDict<TA>& D = M.pat;
    constexpr int NT = TmaTile<TA>::NT;
    if (!D.tma_ok || !ctx.use_tma || mode == MODE_ADD || sizeof(TA) != sizeof(TV)) return false;
    if (M.n_rows < ctx.tma_min_rows) return false;
    // bulk copies need 16-byte aligned global addresses
    if ((reinterpret_cast<uintptr_t>(x) & 15) || (b && (reinterpret_cast<uintptr_t>(b) & 15)) ||
        (d && (reinterpret_cast<uintptr_t>(d) & 15)) || x == y)
        return false;
    const bool need_b = (mode == MODE_RESID || mode == MODE_SWEEP);
    const bool need_d = (mode == MODE_SWEEP && !dpat);
    const size_t head = ((64 + (size_t)D.nent * sizeof(PatEntry<TA>) + (size_t)D.npat * (sizeof(TV) + 4) + 127) / 128) * 128;
    const int elems = D.plan.total + (need_b ? NT : 0) + (need_d ? NT : 0);
    const size_t stage = (((size_t)elems * sizeof(TV) + (size_t)NT * 2) + 127) / 128 * 128;
    const size_t smem = head + 2 * stage;
    if (smem > (size_t)ctx.max_smem_optin - 1024) return false;
    const int ntiles = t1 < 0 ? cdiv(M.n_rows, NT) : t1;
    if (ntiles <= t0) return false;
    int per = (int)std::min<size_t>(2048 / NT, ((size_t)ctx.max_smem_optin + 1024) / (smem + 1024));
    per = std::max(per, 1);
    const int grid = std::min(ntiles - t0, std::max(ctx.sm_count, ctx.sm_count * per - cdiv(reserve_threads, NT)));
#define MGB_TL(MODE, DP)                                                                                       \
    {                                                                                                          \
        auto kern = pat_tma_kernel<TA, TV, MODE, DP, NT>;                                                      \
        MGB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx.max_smem_optin));         \
        kern<<<grid, NT, smem, ctx.stream>>>(D.plan, pp, M.n_rows, t0, ntiles, D.xlo, D.xhi, D.npat, D.nent, D.pid, D.hdr, \
                                             D.ent_s, dpat, x, b, d, y);                                       \
    }
    if (mode == MODE_SPMV) MGB_TL(MODE_SPMV, false)
    else if (mode == MODE_RESID) MGB_TL(MODE_RESID, false)
    else if (dpat) MGB_TL(MODE_SWEEP, true)
    else MGB_TL(MODE_SWEEP, false)
#undef MGB_TL
    MGB_LAUNCH_CHECK();
    return true;
}

// grid-hinted transfer kernels (grid_xfer.cuh): P in mode ADD, R in mode SPMV, one right-hand side; option
// "grid_transfers" (MGB200_GRID_TRANSFERS, default 1) and a hint that was verified at upload
template <typename TA, typename TV>
static bool launch_grid_xfer(Context& ctx, const Csr<TA>& M, int mode, const TV* x, TV* y, bool dry = false,
                             const PutPlan& pp = no_put()) {
    if constexpr (VT<TA>::is_complex) {
        return false;     // P and R are real (SA-AMG.jl:9-10, MGsetup.jl:80-81): no complex instantiation
    } else {
    const GridXfer& X = M.gx;
    if (!X.ok || !X.tab || !M.pat.present || ctx.grid_transfers <= 0 || x == y) return false;
    if (!((X.kind == 1 && mode == MODE_ADD) || (X.kind == 2 && mode == MODE_SPMV))) return false;
    if (dry) return true;      // the caller only asks whether this kernel will run (byte accounting)
    const int width = X.kind == 1 ? X.n[0] : X.N[0];
    if (pp.on && X.kind != 1) return false;
    const long long nlines = X.kind == 1 ? (long long)X.n[1] * X.nk : (long long)X.N[1] * X.nk;
    const int nt = std::min(1024, (width + 31) / 32 * 32);
    const int grid = (int)std::min<long long>(nlines, (long long)ctx.sm_count * std::max(1, 2048 / nt) * 4);
    if (X.kind == 1 && ctx.gxp_quad) {
        // quad form: coarse cell rows (J, K) of the planes this matrix touches; ny cell rows per CTA pass
        const int ntx = std::min(512, (width + 31) / 32 * 32), ny = std::max(1, 256 / ntx);
        const long long nK = ((X.k0 + X.nk - 1) >> 1) - (X.k0 >> 1) + 1, ngroups = (long long)X.N[1] * nK;
        const int g2 = (int)std::min<long long>((ngroups + ny - 1) / ny, (long long)ctx.sm_count * std::max(1, 2048 / (ntx * ny)) * 2);
        const TA* tab = static_cast<const TA*>(X.tab);
        // register budget per thread: 64 (2 CTAs of 512 threads per SM), 40 (3) or 32 (4) - option gxp_quad = 1, 2, 3
        if (ctx.gxp_quad == 2) gxp_quad_kernel<TA, TV, 3><<<g2, dim3(ntx, ny), 0, ctx.stream>>>(X, pp, tab, x, y);
        else if (ctx.gxp_quad == 3) gxp_quad_kernel<TA, TV, 4><<<g2, dim3(ntx, ny), 0, ctx.stream>>>(X, pp, tab, x, y);
        else gxp_quad_kernel<TA, TV, 2><<<g2, dim3(ntx, ny), 0, ctx.stream>>>(X, pp, tab, x, y);
    } else if (X.kind == 1) gxp_kernel<TA, TV><<<grid, nt, 0, ctx.stream>>>(X, pp, static_cast<const TA*>(X.tab), x, y);
    else gxr_kernel<TA, TV><<<grid, nt, 0, ctx.stream>>>(X, static_cast<const TA*>(X.tab), x, y);
    MGB_LAUNCH_CHECK();
    return true;
    }
}

// block variants of the grid-hinted transfer kernels (nrhs > 1, whole grids)
template <typename TA, typename TV>
static bool launch_grid_xfer_mrhs(Context& ctx, const Csr<TA>& M, int mode, const TV* x, TV* y, int m) {
    if constexpr (VT<TA>::is_complex) {
        return false;
    } else {
        const GridXfer& X = M.gx;
        if (!X.ok || !X.tab || ctx.grid_transfers <= 0 || !ctx.use_patterns || x == y || m < 2) return false;
        if (!((X.kind == 1 && mode == MODE_ADD) || (X.kind == 2 && mode == MODE_SPMV))) return false;
        // measured on cfg4 (profiles/r02u_tune_cfg4.log): the restriction gains on every level (706 -> 551 us per cycle),
        // the prolongation loses against the CSR block kernel (500 -> 874 us): option grid_transfers >= 2 switches it on
        if (X.kind == 1 && ctx.grid_transfers < 2) return false;
        const long long nlines = X.kind == 1 ? (long long)X.n[1] * X.nk : (long long)X.N[1] * X.nk;
        const int grid = (int)std::min<long long>(nlines, (long long)ctx.sm_count * 8 * 4);
        if (X.kind == 1) gxp_mrhs_kernel<TA, TV><<<grid, 256, 0, ctx.stream>>>(X, m, static_cast<const TA*>(X.tab), x, y);
        else gxr_mrhs_kernel<TA, TV><<<grid, 256, 0, ctx.stream>>>(X, m, static_cast<const TA*>(X.tab), x, y);
        MGB_LAUNCH_CHECK();
        return true;
    }
}

// box-stencil kernel (box.cuh): box-structured square operators on a 3-D grid, SPMV / RESID / SWEEP, one right-hand
// side, double and complex double.  Variants (rows per thread RZ, base rows per tile NB, stages): ctx.box_variant.
template <typename TV, int RZ, int NB, int STAGES>
static bool launch_box_variant(Context& ctx, const Csr<TV>& M, int mode, const TV* x, const TV* b, const TV* d,
                               const TV* dpat, TV* y, const PutPlan& pp, bool prepare_only, const BoxWait& bw) {
    BoxDict<TV>& X = const_cast<BoxDict<TV>&>(M.box);      // the record cache is filled on first use
    const PatDict<TV>& D = M.pat;
    BoxPlan P;
    box_make_plan<TV>(P, X.shape, RZ, NB, M.n_rows, D.S, D.S2, D.xlo, D.xhi, X.npat, X.p0);
    P.NP = X.NP;
    if (P.ntiles <= 0 || (long long)M.n_rows + 2LL * D.S2 + NB >= (1LL << 31) || (RZ > 1 && M.n_rows % D.S2 != 0)) return false;
    const bool need_b = (mode == MODE_RESID || mode == MODE_SWEEP), need_d = (mode == MODE_SWEEP && !dpat);
    const bool need_pw = (mode == MODE_SWEEP2_FROM_ZERO);
    const size_t smem = box_head_bytes<TV>(P, X.shape) + STAGES * box_stage_bytes<TV>(P, RZ, NB, need_b, need_d, need_pw);
    if (smem > (size_t)ctx.max_smem_optin) return false;
    // the tile records are planned outside stream capture (box_prepare, called before a cycle graph is captured)
    if (!prepare_only && !X.has_records(P, RZ, NB)) {
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        MGB_CUDA(cudaStreamIsCapturing(ctx.stream, &cs));
        if (cs != cudaStreamCaptureStatusNone) return false;
    }
    const unsigned char* recs = X.records(P, RZ, NB);
    P.first_ghost_tile = X.first_ghost_tile;
    if (prepare_only) return true;
#define MGB_BX(SHAPE, MODE, DP)                                                                                      \
    {                                                                                                                 \
        auto kern = box_kernel<TV, SHAPE, MODE, DP, RZ, NB, STAGES>;                                                  \
        MGB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx.max_smem_optin));        \
        int per = 0;                                                                                                  \
        MGB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, kern, NB, smem));                                \
        if (per < 1) return false;                                                                                    \
        /* a kernel that waits for the exchange beside it must leave that kernel room on every SM */                  \
        if (bw.epoch && per * NB > 1536) return false;                                                                \
        const int grid = (int)std::min<long long>(P.ntiles, (long long)ctx.sm_count * per);                           \
        kern<<<grid, NB, smem, ctx.stream>>>(P, X.c0, pp, bw, recs, D.pid, X.ctab, X.dtab, x, b, d, y);                     \
    }
#define MGB_BXS(MODE, DP) { if (X.shape == 7) MGB_BX(7, MODE, DP) else MGB_BX(27, MODE, DP) }
    if (mode == MODE_SPMV) MGB_BXS(MODE_SPMV, false)
    else if (mode == MODE_RESID) MGB_BXS(MODE_RESID, false)
    else if (mode == MODE_SWEEP2_FROM_ZERO) MGB_BXS(MODE_SWEEP2_FROM_ZERO, true)
    else if (dpat) MGB_BXS(MODE_SWEEP, true)
    else MGB_BXS(MODE_SWEEP, false)
#undef MGB_BXS
#undef MGB_BX
    MGB_LAUNCH_CHECK();
    return true;
}
// direct form (box_direct_kernel): no staging, RZ rows per thread one plane apart
template <typename TV, int RZ>
static bool launch_box_direct(Context& ctx, const Csr<TV>& M, int mode, const TV* x, const TV* b, const TV* d,
                              const TV* dpat, TV* y, const PutPlan& pp, bool prepare_only) {
    const BoxDict<TV>& X = M.box;
    const PatDict<TV>& D = M.pat;
    constexpr int NT = 256;
    BoxPlan P;
    box_make_plan<TV>(P, X.shape, RZ, NT, M.n_rows, D.S, D.S2, D.xlo, D.xhi, X.npat, X.p0);
    P.NP = X.NP;
    if ((long long)M.n_rows + 2LL * D.S2 + NT >= (1LL << 31) || (RZ > 1 && M.n_rows % D.S2 != 0)) return false;
    const int ngroups = (P.nplanes + RZ - 1) / RZ;
    if (ngroups > 65535) return false;
    if (prepare_only) return true;
    const size_t smem = ((size_t)X.shape * P.NP + P.NP) * sizeof(TV);
    const int chunks = (P.plane + NT - 1) / NT;
    const int gx = std::max(1, std::min(chunks, ctx.sm_count * 32 / ngroups + 1));
    const dim3 grid(gx, ngroups);
#define MGB_BD(SHAPE, MODE, DP) \
    box_direct_kernel<TV, SHAPE, MODE, DP, RZ, NT><<<grid, NT, smem, ctx.stream>>>(P, X.c0, pp, D.pid, X.ctab, X.dtab, D.pat_off, D.ent, x, b, d, y)
#define MGB_BDS(MODE, DP) { if (X.shape == 7) MGB_BD(7, MODE, DP); else MGB_BD(27, MODE, DP); }
    if (mode == MODE_SPMV) MGB_BDS(MODE_SPMV, false)
    else if (mode == MODE_RESID) MGB_BDS(MODE_RESID, false)
    else if (dpat) MGB_BDS(MODE_SWEEP, true)
    else MGB_BDS(MODE_SWEEP, false)
#undef MGB_BDS
#undef MGB_BD
    MGB_LAUNCH_CHECK();
    return true;
}
// block variant of the box-stencil kernel (box_mrhs_kernel): nrhs > 1, blocks stored RHS-fastest
template <typename TA, typename TV>
static bool launch_box_mrhs(Context& ctx, const Csr<TA>& M, int mode, const TV* x, const TV* b, const TV* d, const TV* dpat,
                            TV* y, int m) {
    if constexpr (std::is_same<TA, TV>::value && (std::is_same<TV, double>::value || std::is_same<TV, cplx>::value)) {
        const BoxDict<TV>& X = M.box;
        const PatDict<TV>& D = M.pat;
        if (!X.ok || !ctx.use_box || !ctx.use_patterns || mode == MODE_ADD || x == y || m < 2) return false;
        if ((long long)M.n_rows + 2LL * D.S2 + 2048 >= (1LL << 31)) return false;
        // marching form (2.5-D tiles through shared memory): whole grids, Float64, at least 8 right-hand sides (a ComplexF64
        // block needs 205 KB of plane buffers and spills at 64 registers: it keeps the direct form)
        if (std::is_same<TV, double>::value && ctx.mrhs_march && m >= 8 && D.xlo == 0 && M.n_rows == M.n_cols && D.S >= 3 && D.S2 >= 3 * D.S && D.S2 % D.S == 0 &&
            M.n_rows % D.S2 == 0 && march_smem_bytes<TV>(X.shape, X.NP) <= (size_t)ctx.max_smem_optin) {
            const int n0 = D.S, n1 = D.S2 / D.S, nz = M.n_rows / D.S2;
            const int tiles = cdiv(n0, MARCH_TX) * cdiv(n1, MARCH_TY);
            const int want = std::max(1, cdiv(4 * ctx.sm_count, tiles));              // z-chunks for ~4 CTAs per SM
            const int nzc = std::max(1, std::min(want, std::max(1, nz / 4)));
            const int zchunk = cdiv(nz, nzc);
            const size_t smem = march_smem_bytes<TV>(X.shape, X.NP);
            const dim3 grd(tiles, cdiv(nz, zchunk), cdiv(m, 32));
#define MGB_MM(SHAPE, MODE, DP)                                                                                               \
    {                                                                                                                         \
        auto kern = box_mrhs_march_kernel<TV, SHAPE, MODE, DP>;                                                               \
        MGB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx.max_smem_optin));                \
        kern<<<grd, MARCH_NW * 32, smem, ctx.stream>>>(X.c0, X.NP, X.p0, m, n0, n1, nz, zchunk, D.pid, X.ctab, X.dtab, x, b, d, y); \
    }
#define MGB_MMS(MODE, DP) { if (X.shape == 7) MGB_MM(7, MODE, DP) else MGB_MM(27, MODE, DP) }
            if (mode == MODE_SPMV) MGB_MMS(MODE_SPMV, false)
            else if (mode == MODE_RESID) MGB_MMS(MODE_RESID, false)
            else if (dpat) MGB_MMS(MODE_SWEEP, true)
            else MGB_MMS(MODE_SWEEP, false)
#undef MGB_MMS
#undef MGB_MM
            MGB_LAUNCH_CHECK();
            return true;
        }
        constexpr int NT = 256;
        BoxPlan P;
        box_make_plan<TV>(P, X.shape, 1, NT, M.n_rows, D.S, D.S2, D.xlo, D.xhi, X.npat, X.p0);
        P.NP = X.NP;
        int mp = 1;
        while (mp < m && mp < 32) mp *= 2;
        const size_t smem = ((size_t)X.shape * P.NP + P.NP) * sizeof(TV);
        const int rpb = NT / mp;
        const int grid = (int)std::min<long long>(((long long)M.n_rows + rpb - 1) / rpb, (long long)ctx.sm_count * 32);
#define MGB_BM(SHAPE, MODE, DP) \
    box_mrhs_kernel<TV, SHAPE, MODE, DP, NT><<<grid, NT, smem, ctx.stream>>>(P, X.c0, m, mp, D.pid, X.ctab, X.dtab, D.pat_off, D.ent, x, b, d, y)
#define MGB_BMS(MODE, DP) { if (X.shape == 7) MGB_BM(7, MODE, DP); else MGB_BM(27, MODE, DP); }
        if (mode == MODE_SPMV) MGB_BMS(MODE_SPMV, false)
        else if (mode == MODE_RESID) MGB_BMS(MODE_RESID, false)
        else if (dpat) MGB_BMS(MODE_SWEEP, true)
        else MGB_BMS(MODE_SWEEP, false)
#undef MGB_BMS
#undef MGB_BM
        MGB_LAUNCH_CHECK();
        return true;
    } else {
        return false;
    }
}

template <typename TA, typename TV>
static bool launch_box(Context& ctx, const Csr<TA>& M, int mode, const TV* x, const TV* b, const TV* d, const TV* dpat,
                       TV* y, const PutPlan& pp, bool prepare_only = false, const BoxWait& bw = no_wait()) {
    if constexpr (std::is_same<TA, TV>::value && (std::is_same<TV, double>::value || std::is_same<TV, cplx>::value)) {
        if (!M.box.ok || !ctx.use_box || M.n_rows < ctx.box_min_rows) return false;
        // Variant (rows per thread RZ, base rows per tile NB, stages).  Measured choices (profiles/r02_box_variants.log,
        // r02t_*): a window carries a y-halo of 2 (S + 1) elements, so long lines need long tiles.
        //   Float64 7-point: (2, 512, 2); lines of >= 400 nodes (the z-slabs of the 512^3 grid) and the fused first two
        //   sweeps (126 -> 95 us at 257^3: fewer neighbours are recomputed per row): (4, 1024, 2)
        //   Float64 27-point: (2, 512, 2)
        //   ComplexF64 7-point: (1, 512, 2); lines of >= 400 nodes (cfg5, 513^3): (2, 1024, 2), the fused first two sweeps
        //   (whose stage also holds the pattern-id windows) (2, 512, 2)
        // Options box_variant / box_variant27 / box_variant_c >= 0 override (tools/tune.py).
        int variant = 1;
        if (sizeof(TV) == 8 && M.box.shape == 7 && (M.pat.S >= 400 || mode == MODE_SWEEP2_FROM_ZERO)) variant = 11;
        if (sizeof(TV) == 16 && M.box.shape == 7) variant = M.pat.S >= 400 ? (mode == MODE_SWEEP2_FROM_ZERO ? 22 : 31) : 21;
        if (ctx.box_variant >= 0) variant = ctx.box_variant;
        if (M.box.shape == 27 && ctx.box_variant27 >= 0) variant = ctx.box_variant27;
        if (sizeof(TV) == 16 && ctx.box_variant_c >= 0) variant = ctx.box_variant_c;      // complex double: its own choice
        if (!prepare_only) {
            if (mode == MODE_ADD || x == y) return false;
            // the fused first two sweeps: b is the staged vector, d must be folded, no ghost rows, staged variants only
            if (mode == MODE_SWEEP2_FROM_ZERO && (!dpat || pp.on || M.pat.xlo != 0 || M.n_rows != M.n_cols || (variant >= 8 && variant <= 10)))
                return false;
            if ((reinterpret_cast<uintptr_t>(x) & 15) || (b && (reinterpret_cast<uintptr_t>(b) & 15)) ||
                (d && (reinterpret_cast<uintptr_t>(d) & 15)))
                return false;
        }
        constexpr int F = sizeof(TV) / 8;      // complex tiles hold half the rows
        switch (variant) {
            case 1: return launch_box_variant<TV, 2, 512 / F, 2>(ctx, M, mode, x, b, d, dpat, y, pp, prepare_only, bw);
            case 3: return launch_box_variant<TV, 4, 256 / F, 1>(ctx, M, mode, x, b, d, dpat, y, pp, prepare_only, bw);
            case 11: return launch_box_variant<TV, 4, 1024 / F, 2>(ctx, M, mode, x, b, d, dpat, y, pp, prepare_only, bw);
            // the same tile sizes for both value types (the variants above halve NB for complex values)
            case 20: return launch_box_variant<TV, 2, 512, 1>(ctx, M, mode, x, b, d, dpat, y, pp, prepare_only, bw);
            case 21: return launch_box_variant<TV, 1, 512, 2>(ctx, M, mode, x, b, d, dpat, y, pp, prepare_only, bw);
            case 22: return launch_box_variant<TV, 2, 512, 2>(ctx, M, mode, x, b, d, dpat, y, pp, prepare_only, bw);
            // long lines (S >= 512: the y-halo of a window, 2 (S + 1) elements, must not dwarf the tile): 1024 base rows
            case 30: return launch_box_variant<TV, 1, 1024, 2>(ctx, M, mode, x, b, d, dpat, y, pp, prepare_only, bw);
            case 31: return launch_box_variant<TV, 2, 1024, 2>(ctx, M, mode, x, b, d, dpat, y, pp, prepare_only, bw);
            case 32: return launch_box_variant<TV, 2, 1024, 1>(ctx, M, mode, x, b, d, dpat, y, pp, prepare_only, bw);
            case 33: return launch_box_variant<TV, 4, 1024, 1>(ctx, M, mode, x, b, d, dpat, y, pp, prepare_only, bw);
            case 9: return bw.epoch ? false : launch_box_direct<TV, 2>(ctx, M, mode, x, b, d, dpat, y, pp, prepare_only);
            default: return launch_box_variant<TV, 1, 1024 / F, 2>(ctx, M, mode, x, b, d, dpat, y, pp, prepare_only, bw);
        }
    } else {
        return false;
    }
}

// plan the tile records of the box-stencil kernel for the variant in use (allocates: never inside stream capture)
template <typename TA>
static void box_prepare(Context& ctx, const Csr<TA>& M) {
    launch_box<TA, TA>(ctx, M, MODE_SPMV, nullptr, nullptr, nullptr, nullptr, nullptr, no_put(), true);
    launch_box<TA, TA>(ctx, M, MODE_SWEEP2_FROM_ZERO, nullptr, nullptr, nullptr, nullptr, nullptr, no_put(), true);   // may run another variant
}

// one-pass dictionary kernel over the rows [rA, rA + nA) and [rB, rB + nB)
template <typename TA, typename TV>
static void launch_pattern_rows(Context& ctx, const Csr<TA>& M, int mode, const TV* x, const TV* b, const TV* d,
                                const TV* dpat, TV* y, int rA, int nA, int rB, int nB, const PutPlan& pp = no_put()) {
    const PatDict<TA>& D = M.pat;
    if (nA + nB <= 0) return;
    const int nt = 256, grid = cdiv(nA + nB, nt);
#define MGB_PL(MODE, RR, DP) \
    pat_kernel<TA, TV, MODE, RR, DP><<<grid, nt, 0, ctx.stream>>>(pp, rA, nA, rB, nB, D.pid, D.c0, D.pat_off, D.ent, dpat, x, b, d, y)
#define MGB_PCASE(MODE)                                     \
    case MODE:                                              \
        if (D.rowrel) {                                     \
            if (MODE == MODE_SWEEP && dpat) MGB_PL(MODE, true, true); \
            else MGB_PL(MODE, true, false);                 \
        } else {                                            \
            if (MODE == MODE_SWEEP && dpat) MGB_PL(MODE, false, true); \
            else MGB_PL(MODE, false, false);                \
        }                                                   \
        break;
    switch (mode) {
        MGB_PCASE(MODE_SPMV)
        MGB_PCASE(MODE_ADD)
        MGB_PCASE(MODE_RESID)
        MGB_PCASE(MODE_SWEEP)
    }
#undef MGB_PCASE
#undef MGB_PL
    MGB_LAUNCH_CHECK();
}

template <typename TA, typename TV>
static void launch_pattern_mode(Context& ctx, const Csr<TA>& M, int mode, const TV* x, const TV* b, const TV* d,
                                const TV* dpat, TV* y, const PutPlan& pp = no_put()) {
    static const bool dbg = env_int("MGB200_DEBUG_KERNELS", 0) != 0;
    if (launch_box<TA, TV>(ctx, M, mode, x, b, d, dpat, y, pp)) {
        if (dbg) std::fprintf(stderr, "[mgb200 dev %d] rows %d mode %d: box kernel (shape %d)\n", ctx.device, M.n_rows, mode, M.box.shape);
        return;
    }
    if (dbg)
        std::fprintf(stderr, "[mgb200 dev %d] rows %d mode %d: dictionary walk (box ok %d, S %d, S2 %d, npat %d, rows %% S2 = %d, x align %d)\n",
                     ctx.device, M.n_rows, mode, (int)M.box.ok, M.pat.S, M.pat.S2, M.pat.npat,
                     M.pat.S2 > 0 ? (int)(M.n_rows % M.pat.S2) : -1, (int)(reinterpret_cast<uintptr_t>(x) & 127));
    if (M.pat.rowrel && launch_pattern_tma<TA, TV>(ctx, M, mode, x, b, d, dpat, y, 0, -1, 0, pp)) return;
    launch_pattern_rows<TA, TV>(ctx, M, mode, x, b, d, dpat, y, 0, M.n_rows, 0, 0, pp);
}

// Split form of a dictionary pass: the rows [lo, hi) first, then `between()` (the caller joins the stream that
// carried the halo exchange of x), then the remaining rows at both ends.  The interior runs as whole tiles of the
// TMA-staged kernel when the matrix qualifies (the ends grow to the tile boundaries).  Same threads-per-row and
// accumulation order as the unsplit pass: bit-identical results.
template <typename TA, typename TV, typename F>
static void pattern_apply_split(Context& ctx, const Csr<TA>& M, int mode, const TV* x, const TV* b, const TV* d,
                                const TV* dpat, TV* y, int lo, int hi, int reserve_threads, F&& between,
                                const PutPlan& pp = no_put()) {
    constexpr int NT = TmaTile<TA>::NT;
    int a0 = lo, a1 = hi;
    bool done = false;
    if (M.pat.rowrel) {
        const int t0 = cdiv(lo, NT), t1 = hi / NT;
        if (t1 > t0 && launch_pattern_tma<TA, TV>(ctx, M, mode, x, b, d, dpat, y, t0, t1, reserve_threads, pp)) {
            a0 = t0 * NT;
            a1 = t1 * NT;
            done = true;
        }
    }
    if (!done) launch_pattern_rows<TA, TV>(ctx, M, mode, x, b, d, dpat, y, a0, a1 - a0, 0, 0, pp);
    between();
    launch_pattern_rows<TA, TV>(ctx, M, mode, x, b, d, dpat, y, 0, a0, a1, M.n_rows - a1, pp);
}

// y = op(M x): the one entry point the cycle uses for A, P and R.  `pp` (dictionary kernels only - the caller checks
// pattern_in_use before it relies on the put) makes the kernel store the slab-end rows of y to the neighbours.
template <typename TA, typename TV>
static void csr_apply(Context& ctx, const Csr<TA>& M, int mode, const TV* x, const TV* b, const TV* d, TV* y,
                      int m, int kind, int level, const TV* dpat = nullptr, const PutPlan& pp = no_put()) {
    MGB_CHECK(M.present(), "matrix not uploaded");
    const bool use_pat = M.pat.present && m == 1 && ctx.use_patterns;
    MGB_CHECK(!pp.on || use_pat, "fused put needs the stencil-dictionary format");
    const bool use_gx = use_pat && ctx.grid_transfers > 0 && launch_grid_xfer<TA, TV>(ctx, M, mode, x, y, true, pp);
    // bytes the device format really streams: the grid-hinted transfer kernels read no matrix stream at all
    const double fmt = use_gx ? vec_bytes<TA, TV>(M, mode, m, false)
                              : (use_pat ? M.pat.matrix_bytes(M.n_rows) + vec_bytes<TA, TV>(M, mode, m, dpat != nullptr) : -1.0);
    Launch L(ctx, kind, level, csr_bytes<TA, TV>(M, mode, m), fmt);
    if (use_gx && launch_grid_xfer<TA, TV>(ctx, M, mode, x, y, false, pp)) return;
    if (use_pat && ctx.split_test > 0 && M.n_rows > 2 * ctx.split_test) {
        // test hook (mgb200_set_option "split_test"): the split launch sequence of the multi-GPU overlap path
        pattern_apply_split<TA, TV>(ctx, M, mode, x, b, d, dpat, y, ctx.split_test, M.n_rows - ctx.split_test, 16384, [] {}, pp);
    } else if (use_pat) {
        launch_pattern_mode<TA, TV>(ctx, M, mode, x, b, d, dpat, y, pp);
    } else if (!M.staged) {
        launch_rowwarp_mode<TA, TV>(ctx, M, mode, x, b, d, y, m);
    } else if (m > 1) {
        // blocks of right-hand sides: the box-stencil kernel when the matrix has that structure, else the CSR stream
        if (!launch_grid_xfer_mrhs<TA, TV>(ctx, M, mode, x, y, m) &&
            !launch_box_mrhs<TA, TV>(ctx, M, mode, x, b, d, ctx.mrhs_dpat ? dpat : nullptr, y, m))
            launch_mrhs_mode<TA, TV>(ctx, M, mode, x, b, d, y, m);
    } else {
        switch (M.tpr) {
            case 1: launch_stream_mode<TA, TV, 1>(ctx, M, mode, x, b, d, y); break;
            case 2: launch_stream_mode<TA, TV, 2>(ctx, M, mode, x, b, d, y); break;
            case 4: launch_stream_mode<TA, TV, 4>(ctx, M, mode, x, b, d, y); break;
            case 8: launch_stream_mode<TA, TV, 8>(ctx, M, mode, x, b, d, y); break;
            case 16: launch_stream_mode<TA, TV, 16>(ctx, M, mode, x, b, d, y); break;
            default: launch_stream_mode<TA, TV, 32>(ctx, M, mode, x, b, d, y); break;
        }
    }
}

// y = op(M x) with the rows [lo, hi) launched first and `between()` called before the rest (dictionary matrices,
// one right-hand side: the caller checks pattern_in_use)
template <typename TA>
static bool pattern_in_use(const Context& ctx, const Csr<TA>& M, int m) {
    return M.pat.present && m == 1 && ctx.use_patterns;
}
template <typename TA, typename TV, typename F>
static void csr_apply_split(Context& ctx, const Csr<TA>& M, int mode, const TV* x, const TV* b, const TV* d, TV* y,
                            int kind, int level, const TV* dpat, int lo, int hi, int reserve_threads, F&& between,
                            const PutPlan& pp = no_put()) {
    MGB_CHECK(pattern_in_use(ctx, M, 1), "split launch needs the stencil-dictionary format");
    const double fmt = M.pat.matrix_bytes(M.n_rows) + vec_bytes<TA, TV>(M, mode, 1, dpat != nullptr);
    Launch L(ctx, kind, level, csr_bytes<TA, TV>(M, mode, 1), fmt);
    pattern_apply_split<TA, TV>(ctx, M, mode, x, b, d, dpat, y, lo, hi, reserve_threads, between, pp);
}

// ---- reductions / vector ops -----------------------------------------------------------------
template <typename TV>
static void dev_dot(Context& ctx, long long n, const TV* x, const TV* y, double* out) {
    Launch L(ctx, K_REDUCE, 0, 2.0 * n * sizeof(TV));
    dot_kernel<TV><<<ctx.red_blocks(n), RED_THREADS, 0, ctx.stream>>>(n, x, y, ctx.red, out);
    MGB_LAUNCH_CHECK();
}
template <typename TV>
static void dev_norm2sq(Context& ctx, long long n, const TV* x, double* out) {
    Launch L(ctx, K_REDUCE, 0, 1.0 * n * sizeof(TV));
    norm2sq_kernel<TV><<<ctx.red_blocks(n), RED_THREADS, 0, ctx.stream>>>(n, x, ctx.red, out);
    MGB_LAUNCH_CHECK();
}
// t[c] = <V[:,c], w>, c < k  -> out (2 doubles per column)
template <typename TV>
static void dev_multi_dot(Context& ctx, long long n, const TV* V, long long ld, int k, const TV* w, double* out) {
    for (int c0 = 0; c0 < k; c0 += 8) {
        int kk = std::min(8, k - c0);
        Launch L(ctx, K_REDUCE, 0, (1.0 + kk) * n * sizeof(TV));
        if (kk <= 2)
            multi_dot_kernel<TV, 2><<<ctx.red_blocks(n), RED_THREADS, 0, ctx.stream>>>(n, V + c0 * ld, ld, kk, w, ctx.red, out + 2 * c0);
        else if (kk <= 4)
            multi_dot_kernel<TV, 4><<<ctx.red_blocks(n), RED_THREADS, 0, ctx.stream>>>(n, V + c0 * ld, ld, kk, w, ctx.red, out + 2 * c0);
        else
            multi_dot_kernel<TV, 8><<<ctx.red_blocks(n), RED_THREADS, 0, ctx.stream>>>(n, V + c0 * ld, ld, kk, w, ctx.red, out + 2 * c0);
        MGB_LAUNCH_CHECK();
    }
}
// w = beta*w + sum_c coef[c] V[:,c]; if norm_out: *norm_out = ||w||^2
template <typename TV>
static void dev_multi_axpy(Context& ctx, long long n, const TV* V, long long ld, int k, const TV* coef,
                           double beta, TV* w, double* norm_out) {
    MGB_CHECK(k <= MAXK, "too many basis columns");
    Coefs<TV> c;
    for (int i = 0; i < MAXK; ++i) c.c[i] = (i < k) ? coef[i] : VT<TV>::zero();
    Launch L(ctx, K_VECTOR, 0, (2.0 + k) * n * sizeof(TV));
    if (norm_out)
        multi_axpy_kernel<TV, true><<<ctx.red_blocks(n), RED_THREADS, 0, ctx.stream>>>(n, V, ld, k, c, beta, w, ctx.red, norm_out);
    else
        multi_axpy_kernel<TV, false><<<ctx.red_blocks(n), RED_THREADS, 0, ctx.stream>>>(n, V, ld, k, c, beta, w, ctx.red, nullptr);
    MGB_LAUNCH_CHECK();
}
template <typename TV>
static void dev_axpby(Context& ctx, long long n, TV a, const TV* x, TV b, TV* y, bool b_is_zero) {
    Launch L(ctx, K_VECTOR, 0, (b_is_zero ? 2.0 : 3.0) * n * sizeof(TV));
    axpby_kernel<TV><<<ctx.ew_blocks(n), 256, 0, ctx.stream>>>(n, a, x, b, y, b_is_zero ? 1 : 0);
    MGB_LAUNCH_CHECK();
}
template <typename TV>
static void dev_copy(Context& ctx, long long n, const TV* src, TV* dst) {
    if (src == dst) return;
    Launch L(ctx, K_COPY, 0, 2.0 * n * sizeof(TV));
    MGB_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(TV), cudaMemcpyDeviceToDevice, ctx.stream));
}
template <typename TV>
static void dev_zero(Context& ctx, long long n, TV* dst) {
    Launch L(ctx, K_COPY, 0, 1.0 * n * sizeof(TV));
    MGB_CUDA(cudaMemsetAsync(dst, 0, n * sizeof(TV), ctx.stream));
}
// read k doubles from a device scalar array (synchronises the stream)
static inline void read_scalars(Context& ctx, const double* dptr, int k, double* out) {
    MGB_CUDA(cudaMemcpyAsync(ctx.scal_host, dptr, k * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
    ctx.sync();
    std::memcpy(out, ctx.scal_host, k * sizeof(double));
}

}  // namespace mgb200
