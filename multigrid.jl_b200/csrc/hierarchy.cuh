// Device-resident multigrid hierarchy: storage, kernel selection, launch wrappers.
#pragma once
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "common.cuh"
#include "csr_kernels.cuh"
#include "dense_lu.cuh"
#include "pattern.cuh"
#include "box.cuh"
#include "grid_xfer.cuh"
#include "galerkin.cuh"
#include "vec_kernels.cuh"

namespace mgb200 {

enum Kind {
    K_SWEEP = 0, K_RESID = 1, K_SPMV = 2, K_RESTRICT = 3, K_PROLONG = 4, K_DIAG = 5, K_COARSE = 6,
    K_REDUCE = 7, K_VECTOR = 8, K_COPY = 9, K_FIRST2 = 10 /* first two sweeps from x = 0, fused */, K_TAIL = 11 /* coarse levels in one kernel */,
    K_NKINDS = 12
};

static inline int env_int(const char* name, int dflt) {
    const char* s = std::getenv(name);
    return s ? std::atoi(s) : dflt;
}

template <typename T>
static T* dev_alloc(size_t n) {
    T* p = nullptr;
    MGB_CUDA(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
    return p;
}
// Vectors of a row-partitioned level keep their lower ghost rows IN FRONT of the owned rows (dist.cuh): the
// pointer handed out is base + pad.  The registry lets dev_free release such a pointer like any other.
struct PaddedRegistry {
    std::mutex mu;
    std::unordered_map<const void*, void*> base;
};
static PaddedRegistry& padded_registry() {
    static PaddedRegistry r;
    return r;
}
// Every vector carries 4 elements of zeroed slack and starts on a 16-byte boundary (pad rounded up to a multiple of
// tma_align<T>() elements: 2 for the 8- and 16-byte value types, 4 for Float32), so that the bulk copies of the TMA
// kernels may round their ranges outwards to that granularity.
template <typename T>
constexpr int tma_align() { return sizeof(T) >= 8 ? 2 : 4; }
// The pad in front of the owned rows (the lower ghost rows of a row-partitioned level) is rounded up to 16 elements, so
// that the owned rows start on a 128-byte boundary like the vectors of a single-GPU run: with a pad that was only
// 16-byte aligned the rank with lower ghosts ran its level-1 sweeps at 150 us against 96 us on the rank without
// (profiles/r02l_bench_n2.log: every 256-byte row segment of a warp straddled three 128-byte lines instead of two).
template <typename T>
static inline size_t align_pad(size_t pad) { return (pad + 15) & ~(size_t)15; }
template <typename T>
static T* vec_alloc(size_t total, size_t pad, cudaStream_t stream) {
    T* b = dev_alloc<T>(total + 4 + 16);
    MGB_CUDA(cudaMemsetAsync(b, 0, (total + 4 + 16) * sizeof(T), stream));
    pad = align_pad<T>(pad);
    if (pad == 0) return b;
    PaddedRegistry& r = padded_registry();
    std::lock_guard<std::mutex> g(r.mu);
    r.base[b + pad] = b;
    return b + pad;
}
template <typename T>
static void dev_free(T*& p) {
    if (p) {
        void* b = p;
        {
            PaddedRegistry& r = padded_registry();
            std::lock_guard<std::mutex> g(r.mu);
            auto it = r.base.find(p);
            if (it != r.base.end()) {
                b = it->second;
                r.base.erase(it);
            }
        }
        cudaFree(b);
    }
    p = nullptr;
}

// ---------------------------------------------------------------------------------------------
// CSR matrix on the device + the kernel configuration chosen from its row-length statistics
// ---------------------------------------------------------------------------------------------
template <typename TA>
struct Csr {
    int n_rows = 0, n_cols = 0;
    long long nnz = 0;
    int* rowptr = nullptr;
    int* colind = nullptr;
    TA* val = nullptr;
    int tpr = 1;        // threads per row
    int rpc = 256;      // rows per CTA
    int cap = 0;        // staged elements per CTA (multiple of 4)
    int max_len = 0;
    bool staged = true; // TMA-staged kernel, else row-per-warp fallback
    size_t smem = 0;
    PatDict<TA> pat;    // stencil-dictionary form (pattern.cuh), when the rows deduplicate
    int int_lo = 0, int_hi = 0;  // row-partitioned levels: rows [int_lo, int_hi) read no ghost row of the input vector
    GridXfer gx{};      // grid hint of a transfer operator (grid_xfer.cuh), ok = 0 unless verified at upload
    BoxDict<TA> box;    // dense coefficient tables of a box-structured dictionary (box.cuh)
    bool present() const { return rowptr != nullptr; }
    void release() {
        pat.release();
        box.release();
        if (gx.tab) cudaFree(gx.tab);
        gx = no_grid();
        int_lo = int_hi = 0;
        dev_free(rowptr);
        dev_free(colind);
        dev_free(val);
        n_rows = n_cols = 0;
        nnz = 0;
    }
};

struct ProfRec {
    int kind, level;
    double bytes;      // algorithmic bytes (SURVEY 8(d) accounting)
    double fmt_bytes;  // bytes of the device format actually streamed (== bytes for plain CSR)
    cudaEvent_t e0, e1;
};

struct Context {
    int device = 0;
    cudaStream_t stream = nullptr;
    ReduceWs red{nullptr, nullptr};
    double* scal = nullptr;        // device scalars (64 doubles)
    double* scal_host = nullptr;   // pinned mirror
    int smem_budget = 56 * 1024;
    int use_patterns = 1;          // 0: always stream CSR (MGB200_PATTERNS / mgb200_set_option)
    int use_graphs = 1;            // 0: never replay cycles from CUDA graphs
    int use_tma = 1;               // 0: never use the TMA-staged dictionary kernel (MGB200_TMA)
    int grid_transfers = 1;        // > 0: grid-hinted transfer kernels (grid_xfer.cuh) where a hint was given and verified
    int gxp_quad = 1;              // prolongation: quad form (four fine lines per thread; 1 / 2 / 3: register budget 64 / 40 / 32)
                                   // instead of the line form (MGB200_GXP_QUAD)
    int tma_min_rows = 200000;     // smaller matrices keep the one-pass kernel (too few tiles per SM)
    int use_box = 1;               // box-stencil kernel (box.cuh) for box-structured square operators (MGB200_BOX)
    int box_variant = -1;          // >= 0: (rows per thread, base rows per tile, stages) instead of launch_box's own choice (MGB200_BOX_VARIANT)
    int box_variant27 = -1;        // >= 0: another variant for the 27-point levels (MGB200_BOX_VARIANT27)
    int box_variant_c = -1;        // >= 0: another variant for ComplexF64 hierarchies (MGB200_BOX_VARIANT_C)
    int box_min_rows = 100000;
    int overlap_box = 1;           // row-partitioned levels: the box kernel runs beside the halo exchange of its input
                                   // vector and waits for the ghost rows inside the kernel (MGB200_OVERLAP_BOX)
    int mrhs_march = 1;            // block variant of the box kernel: marching (2.5-D) form (MGB200_MRHS_MARCH)
    int mrhs_dpat = 0;             // the block kernels take d as a vector (the cycle passes dpat only for one RHS)
    int fuse_first_sweeps = 1;     // first two sweeps from x = 0 in one pass of the box kernel (MGB200_FUSE_FIRST)
    int split_test = 0;            // > 0: every dictionary pass runs as interior + both ends (test hook)
    int use_overlap = 0;           // multi-GPU: halo exchange beside the interior rows (MGB200_OVERLAP=1; measured
                                   // slower than the serial exchange at N = 2, profiles/r01e_bench_n2_*: default off)
    cudaStream_t side = nullptr;   // carries the halo exchange of an overlapped pass
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int max_smem_optin = 0;
    int sm_count = 148;
    bool profiling = false;
    std::vector<ProfRec> prof;
    std::vector<cudaEvent_t> ev_pool;
    std::vector<cudaEvent_t> user_ev;
    long long launches = 0;

    void init(int dev) {
        device = dev;
        MGB_CUDA(cudaSetDevice(dev));
        MGB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        red.partials = dev_alloc<double>((size_t)RED_MAX_BLOCKS * 2 * (MAXK + 2));
        red.counter = dev_alloc<unsigned>(1);
        MGB_CUDA(cudaMemset(red.counter, 0, sizeof(unsigned)));
        scal = dev_alloc<double>(256);
        MGB_CUDA(cudaMemset(scal, 0, 256 * sizeof(double)));
        MGB_CUDA(cudaMallocHost(&scal_host, 256 * sizeof(double)));
        MGB_CUDA(cudaDeviceGetAttribute(&max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        MGB_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
        smem_budget = env_int("MGB200_SMEM_BUDGET", 56 * 1024);
        use_patterns = env_int("MGB200_PATTERNS", 1);
        use_graphs = env_int("MGB200_GRAPHS", 1);
        use_tma = env_int("MGB200_TMA", 1);
        tma_min_rows = env_int("MGB200_TMA_MIN_ROWS", 200000);
        use_box = env_int("MGB200_BOX", 1);
        box_variant = env_int("MGB200_BOX_VARIANT", -1);
        box_variant27 = env_int("MGB200_BOX_VARIANT27", -1);
        box_variant_c = env_int("MGB200_BOX_VARIANT_C", -1);
        box_min_rows = env_int("MGB200_BOX_MIN_ROWS", 100000);
        fuse_first_sweeps = env_int("MGB200_FUSE_FIRST", 1);
        overlap_box = env_int("MGB200_OVERLAP_BOX", 1);
        grid_transfers = env_int("MGB200_GRID_TRANSFERS", 1);
        gxp_quad = env_int("MGB200_GXP_QUAD", 1);
        mrhs_march = env_int("MGB200_MRHS_MARCH", 1);
        use_overlap = env_int("MGB200_OVERLAP", 0);
        split_test = env_int("MGB200_SPLIT_TEST", 0);
        int prio_lo = 0, prio_hi = 0;
        MGB_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        MGB_CUDA(cudaStreamCreateWithPriority(&side, cudaStreamNonBlocking, prio_hi));
        MGB_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
        MGB_CUDA(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
    }
    void destroy() {
        if (!stream) return;
        cudaSetDevice(device);
        cudaStreamSynchronize(stream);
        for (auto& r : prof) {
            cudaEventDestroy(r.e0);
            cudaEventDestroy(r.e1);
        }
        prof.clear();
        for (auto e : ev_pool) cudaEventDestroy(e);
        ev_pool.clear();
        for (auto e : user_ev) cudaEventDestroy(e);
        user_ev.clear();
        dev_free(red.partials);
        dev_free(red.counter);
        dev_free(scal);
        if (scal_host) cudaFreeHost(scal_host);
        scal_host = nullptr;
        if (side) {
            cudaStreamSynchronize(side);
            cudaStreamDestroy(side);
        }
        side = nullptr;
        if (ev_fork) cudaEventDestroy(ev_fork);
        if (ev_join) cudaEventDestroy(ev_join);
        ev_fork = ev_join = nullptr;
        cudaStreamDestroy(stream);
        stream = nullptr;
    }
    cudaEvent_t get_event() {
        if (!ev_pool.empty()) {
            cudaEvent_t e = ev_pool.back();
            ev_pool.pop_back();
            return e;
        }
        cudaEvent_t e;
        MGB_CUDA(cudaEventCreate(&e));
        return e;
    }
    void sync() { MGB_CUDA(cudaStreamSynchronize(stream)); }
    int red_blocks(long long n) const {
        long long b = (n + RED_THREADS * 4 - 1) / (RED_THREADS * 4);
        return (int)std::max<long long>(1, std::min<long long>(b, std::min(RED_MAX_BLOCKS, sm_count * 8)));
    }
    int ew_blocks(long long n) const {
        long long b = (n + 255) / 256;
        return (int)std::max<long long>(1, std::min<long long>(b, (long long)sm_count * 16));
    }
};

// RAII launch bracket: counts the launch and, when profiling, times it with CUDA events on the
// launching stream.
struct Launch {
    Context& c;
    bool rec;
    ProfRec r;
    Launch(Context& ctx, int kind, int level, double bytes, double fmt_bytes = -1.0) : c(ctx), rec(ctx.profiling) {
        c.launches++;
        if (rec) {
            r.kind = kind;
            r.level = level;
            r.bytes = bytes;
            r.fmt_bytes = fmt_bytes < 0 ? bytes : fmt_bytes;
            r.e0 = c.get_event();
            r.e1 = c.get_event();
            cudaEventRecord(r.e0, c.stream);
        }
    }
    void cancel() {          // nothing was launched under this bracket after all
        c.launches--;
        if (rec) {
            c.ev_pool.push_back(r.e0);
            c.ev_pool.push_back(r.e1);
        }
        rec = false;
    }
    ~Launch() {
        if (rec) {
            cudaEventRecord(r.e1, c.stream);
            c.prof.push_back(r);
        }
    }
};

#define MGB_LAUNCH_CHECK() MGB_CUDA(cudaGetLastError())

// ---------------------------------------------------------------------------------------------
// upload: host CSC-of-adjoint (Int64, index_base) -> device CSR-of-operator (Int32)
// ---------------------------------------------------------------------------------------------
template <typename TA>
static void upload_csr(Context& ctx, Csr<TA>& M, long long n_rows, long long n_cols, const int64_t* colptr,
                       const int64_t* rowval, const TA* nzval, int base, bool conjugate,
                       bool want_patterns = true, HostPatterns<TA>* keep = nullptr) {
    MGB_CHECK(n_rows > 0 && n_rows < (1LL << 31) - 8, "matrix rows out of range");
    MGB_CHECK(colptr && rowval && nzval, "null matrix array");
    const long long nnz = colptr[n_rows] - base;
    MGB_CHECK(nnz >= 0 && nnz < (1LL << 31) - 64, "nnz does not fit 32-bit device indices");
    M.release();
    M.n_rows = (int)n_rows;
    M.n_cols = (int)n_cols;
    M.nnz = nnz;
    M.rowptr = dev_alloc<int>(n_rows + 1);
    M.colind = dev_alloc<int>(nnz + 16);
    M.val = dev_alloc<TA>(nnz + 16);
    MGB_CUDA(cudaMemsetAsync(M.colind + nnz, 0, 16 * sizeof(int), ctx.stream));
    MGB_CUDA(cudaMemsetAsync(M.val + nnz, 0, 16 * sizeof(TA), ctx.stream));
    const long long chunk = 32LL << 20;  // elements per staging pass
    long long* tmp = dev_alloc<long long>(std::min<long long>(chunk, std::max(nnz, n_rows + 1)));
    // row pointers
    for (long long o = 0; o < n_rows + 1; o += chunk) {
        long long c = std::min(chunk, n_rows + 1 - o);
        MGB_CUDA(cudaMemcpyAsync(tmp, colptr + o, c * sizeof(long long), cudaMemcpyHostToDevice, ctx.stream));
        convert_index_kernel<<<ctx.ew_blocks(c), 256, 0, ctx.stream>>>(tmp, M.rowptr + o, c, base);
        MGB_LAUNCH_CHECK();
    }
    for (long long o = 0; o < nnz; o += chunk) {
        long long c = std::min(chunk, nnz - o);
        MGB_CUDA(cudaMemcpyAsync(tmp, rowval + o, c * sizeof(long long), cudaMemcpyHostToDevice, ctx.stream));
        convert_index_kernel<<<ctx.ew_blocks(c), 256, 0, ctx.stream>>>(tmp, M.colind + o, c, base);
        MGB_LAUNCH_CHECK();
    }
    ctx.sync();
    dev_free(tmp);
    TA* vtmp = dev_alloc<TA>(std::min<long long>(chunk, std::max<long long>(nnz, 1)));
    for (long long o = 0; o < nnz; o += chunk) {
        long long c = std::min(chunk, nnz - o);
        MGB_CUDA(cudaMemcpyAsync(vtmp, nzval + o, c * sizeof(TA), cudaMemcpyHostToDevice, ctx.stream));
        conj_copy_kernel<TA><<<ctx.ew_blocks(c), 256, 0, ctx.stream>>>(vtmp, M.val + o, c, conjugate ? 1 : 0);
        MGB_LAUNCH_CHECK();
    }
    ctx.sync();
    dev_free(vtmp);

    // ---- kernel selection from row-length statistics --------------------------------------
    int* dstat = dev_alloc<int>(2);
    MGB_CUDA(cudaMemsetAsync(dstat, 0, 2 * sizeof(int), ctx.stream));
    max_rowlen_kernel<<<cdiv(n_rows, 256), 256, 0, ctx.stream>>>(M.rowptr, M.n_rows, dstat);
    MGB_LAUNCH_CHECK();
    int hstat[2];
    MGB_CUDA(cudaMemcpyAsync(hstat, dstat, sizeof(int), cudaMemcpyDeviceToHost, ctx.stream));
    ctx.sync();
    M.max_len = hstat[0];
    const double mean = (double)nnz / (double)n_rows;
    int tpr = 1;
    while (tpr < 32 && mean > 40.0 * tpr) tpr *= 2;
    tpr = env_int("MGB200_TPR", tpr);
    int nt = env_int("MGB200_NT", 256);
    M.tpr = tpr;
    M.staged = false;
    for (; nt >= 32 && nt >= tpr; nt /= 2) {
        int rpc = nt / tpr;
        MGB_CUDA(cudaMemsetAsync(dstat + 1, 0, sizeof(int), ctx.stream));
        int nchunks = cdiv(n_rows, rpc);
        chunk_span_kernel<<<cdiv(nchunks, 256), 256, 0, ctx.stream>>>(M.rowptr, M.n_rows, rpc, dstat + 1);
        MGB_LAUNCH_CHECK();
        MGB_CUDA(cudaMemcpyAsync(hstat + 1, dstat + 1, sizeof(int), cudaMemcpyDeviceToHost, ctx.stream));
        ctx.sync();
        size_t bytes = 16 + (size_t)hstat[1] * (sizeof(TA) + sizeof(int));
        if (bytes <= (size_t)ctx.smem_budget || (nt / 2 < 32 || nt / 2 < tpr)) {
            if (bytes <= (size_t)ctx.max_smem_optin - 1024) {
                M.rpc = rpc;
                M.cap = hstat[1];
                M.smem = bytes;
                M.staged = true;
            }
            break;
        }
    }
    dev_free(dstat);

    // ---- stencil dictionary (pattern.cuh): deduplicate the rows on the host ---------------------
    if (want_patterns && ctx.use_patterns && n_cols < (1LL << 31) - 8) {
        HostPatterns<TA> hp;
        if (build_patterns<TA>(n_rows, colptr, rowval, nzval, base, conjugate, PAT_MAX_PATTERNS, PAT_MAX_ENTRIES, hp)) {
            upload_patterns<TA>(M.pat, hp, n_rows);
            M.pat.xlo = 0;                               // input vectors hold n_cols elements (+ slack, vec_alloc)
            M.pat.xhi = (long long)align_pad<TA>((size_t)n_cols);
            if (hp.rowrel && ctx.use_box) {     // box-stencil kernel (box.cuh): dense tables
                const BoxInfo B = detect_box<TA>(hp, n_rows);
                BoxDict<TA>& X = M.box;
                if (box_build_tables<TA>(hp, B, n_rows, X.shape, X.NP, X.p0, X.h_ctab, X.c0)) {
                    X.npat = hp.npat();
                    X.h_mask = B.mask;
                    X.h_pat_off = hp.pat_off;
                    X.ctab = dev_alloc<TA>(X.h_ctab.size());
                    MGB_CUDA(cudaMemcpy(X.ctab, X.h_ctab.data(), X.h_ctab.size() * sizeof(TA), cudaMemcpyHostToDevice));
                    X.dtab = dev_alloc<TA>(X.NP);
                    MGB_CUDA(cudaMemset(X.dtab, 0, X.NP * sizeof(TA)));
                    X.ok = true;
                }
            }
            if (keep) *keep = std::move(hp);      // the caller checks a grid hint against the dictionary
        }
    }
}

}  // namespace mgb200
