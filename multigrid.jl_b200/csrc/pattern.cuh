// Stencil-dictionary ("pattern") storage of a CSR matrix and its row-gather kernel.
//
// The geometric hierarchies of the reference (MGsetup.jl:7-138: P = kron of 1-D interpolations,
// A_c = R A P or rediscretisation on a regular mesh) hand us CSR matrices whose rows are copies of
// a handful of stencils: an interior row and its boundary variants.  Streaming 12-20 bytes per
// non-zero for such a matrix is the dominant HBM traffic of the cycle, and all of
// it is redundant.  At upload the host looks for that redundancy in the CSR arrays themselves -- no
// mesh information is passed through the ABI -- by deduplicating rows on the exact byte image of
// (column offsets, values):
//
//     row i  ->  pid[i]  (uint16)  [+ c0[i] (int32) when offsets are taken from the first column]
//     pattern p -> entries ent[pat_off[p] .. pat_off[p+1]) = (value, column offset), stored order
//
// The kernel then reads 2 (or 6) bytes per ROW instead of 12-20 bytes per NON-ZERO; the dictionary
// (a few KB) sits in L1 and is read with warp-uniform broadcast loads.  One thread owns one row and
// accumulates the products in stored order, so results are bit-identical to the CSR kernel and to
// the CPU oracle.  Matrices without such structure (SA-AMG levels, variable coefficients) exceed
// the pattern cap within the first few thousand rows and keep the CSR-stream kernel.
//
// If relaxPrecs[l] (the d of  x += d.*r, MGcycle.jl:129) is also a function of the pattern id it is
// folded into the dictionary as well ("dpat").
#pragma once
#include <atomic>
#include <thread>
#include <unordered_map>
#include <vector>

#include "common.cuh"

namespace mgb200 {

template <typename TA>
struct __align__(16) PatEntry {
    TA v;
    int delta;
};

constexpr int PAT_MAX_PATTERNS = 4096;
constexpr int PAT_MAX_ENTRIES = 1 << 16;
constexpr int PAT_LEN_SHIFT = 20;               // device header of a pattern = entry offset | row length << 20
constexpr int PAT_MAX_LEN = (1 << 11) - 1;

// host result of the row deduplication
template <typename TA>
struct HostPatterns {
    bool ok = false;
    bool rowrel = false;             // column = row + delta, else column = c0[row] + delta
    std::vector<uint16_t> pid;       // n_rows
    std::vector<int> c0;             // n_rows (only when !rowrel)
    std::vector<int> pat_off;        // npat + 1
    std::vector<int> delta;          // entries
    std::vector<TA> val;             // entries (operator values: conjugated if requested)
    std::vector<long long> rep_row;  // representative row of each pattern
    int npat() const { return (int)pat_off.size() - 1; }
};

static inline uint64_t pat_mix(uint64_t h, uint64_t v) {
    h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
    h *= 0xff51afd7ed558ccdull;
    h ^= h >> 32;
    return h;
}

// One pass of the deduplication for a fixed reference mode.  cp/rv/nz are CSR arrays (row
// pointers with index base `base`, Int64 columns, values) of the stored matrix.
template <typename TA>
static bool build_patterns_mode(long long n_rows, const int64_t* cp, const int64_t* rv, const TA* nz, int base,
                                bool conjugate, bool rowrel, int max_pat, int max_ent, HostPatterns<TA>& out) {
    static_assert(sizeof(TA) % 8 == 0, "value type is made of doubles");
    constexpr int W = sizeof(TA) / 8;
    int T = (int)std::thread::hardware_concurrency();
    T = std::max(1, std::min(T, 16));
    if (n_rows < 4096) T = 1;
    struct Local {
        std::unordered_map<uint64_t, std::vector<int>> map;
        std::vector<long long> rep;   // representative row per local pattern
        std::vector<uint64_t> hash;
    };
    std::vector<Local> loc(T);
    std::vector<uint16_t> pid(n_rows);
    std::atomic<bool> fail(false);
    auto ref_of = [&](long long row, long long k0, long long len) -> long long {
        if (rowrel) return row;
        return len > 0 ? rv[k0] - base : 0;
    };
    auto same = [&](long long ra, long long rb) -> bool {
        const long long a0 = cp[ra] - base, a1 = cp[ra + 1] - base, b0 = cp[rb] - base, b1 = cp[rb + 1] - base;
        if (a1 - a0 != b1 - b0) return false;
        const long long refa = ref_of(ra, a0, a1 - a0), refb = ref_of(rb, b0, b1 - b0);
        for (long long k = 0; k < a1 - a0; ++k)
            if (rv[a0 + k] - refa != rv[b0 + k] - refb) return false;
        return std::memcmp(nz + a0, nz + b0, (size_t)(a1 - a0) * sizeof(TA)) == 0;
    };
    auto work = [&](int t) {
        const long long r0 = n_rows * t / T, r1 = n_rows * (t + 1) / T;
        Local& L = loc[t];
        for (long long row = r0; row < r1; ++row) {
            if ((row & 1023) == 0 && fail.load(std::memory_order_relaxed)) return;
            const long long k0 = cp[row] - base, k1 = cp[row + 1] - base;
            const long long ref = ref_of(row, k0, k1 - k0);
            uint64_t h = pat_mix(0x1234567ull, (uint64_t)(k1 - k0));
            const uint64_t* vb = reinterpret_cast<const uint64_t*>(nz + k0);
            for (long long k = k0; k < k1; ++k) {
                h = pat_mix(h, (uint64_t)(rv[k] - base - ref));
                for (int w = 0; w < W; ++w) h = pat_mix(h, vb[(k - k0) * W + w]);
            }
            std::vector<int>& cand = L.map[h];
            int id = -1;
            for (int c : cand)
                if (same(L.rep[c], row)) {
                    id = c;
                    break;
                }
            if (id < 0) {
                id = (int)L.rep.size();
                if (id >= max_pat) {
                    fail.store(true);
                    return;
                }
                L.rep.push_back(row);
                L.hash.push_back(h);
                cand.push_back(id);
            }
            pid[row] = (uint16_t)id;
        }
    };
    if (T == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t) th.emplace_back(work, t);
        for (auto& x : th) x.join();
    }
    if (fail.load()) return false;
    // merge the per-thread dictionaries in (thread, local id) order == order of first appearance
    std::unordered_map<uint64_t, std::vector<int>> gmap;
    std::vector<long long> grep;
    std::vector<std::vector<int>> remap(T);
    for (int t = 0; t < T; ++t) {
        remap[t].resize(loc[t].rep.size());
        for (size_t c = 0; c < loc[t].rep.size(); ++c) {
            std::vector<int>& cand = gmap[loc[t].hash[c]];
            int id = -1;
            for (int g : cand)
                if (same(grep[g], loc[t].rep[c])) {
                    id = g;
                    break;
                }
            if (id < 0) {
                id = (int)grep.size();
                if (id >= max_pat) return false;
                grep.push_back(loc[t].rep[c]);
                cand.push_back(id);
            }
            remap[t][c] = id;
        }
    }
    long long entries = 0;
    for (long long r : grep) {
        entries += cp[r + 1] - cp[r];
        if (cp[r + 1] - cp[r] > PAT_MAX_LEN) return false;
    }
    if (entries > max_ent) return false;
    // worth it only if the dictionary is much smaller than the matrix it replaces
    if ((long long)grep.size() * 8 > n_rows || entries * 4 > cp[n_rows] - base) return false;
    auto fix = [&](int t) {
        const long long r0 = n_rows * t / T, r1 = n_rows * (t + 1) / T;
        for (long long row = r0; row < r1; ++row) pid[row] = (uint16_t)remap[t][pid[row]];
    };
    if (T == 1) {
        fix(0);
    } else {
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t) th.emplace_back(fix, t);
        for (auto& x : th) x.join();
    }
    out.ok = true;
    out.rowrel = rowrel;
    out.pid.swap(pid);
    out.rep_row = grep;
    out.pat_off.assign(1, 0);
    out.delta.clear();
    out.val.clear();
    for (long long r : grep) {
        const long long k0 = cp[r] - base, k1 = cp[r + 1] - base;
        const long long ref = ref_of(r, k0, k1 - k0);
        for (long long k = k0; k < k1; ++k) {
            out.delta.push_back((int)(rv[k] - base - ref));
            out.val.push_back(conjugate ? conj_(nz[k]) : nz[k]);
        }
        out.pat_off.push_back((int)out.delta.size());
    }
    out.c0.clear();
    if (!rowrel) {
        out.c0.resize(n_rows);
        for (long long row = 0; row < n_rows; ++row) {
            const long long k0 = cp[row] - base, k1 = cp[row + 1] - base;
            out.c0[row] = k1 > k0 ? (int)(rv[k0] - base) : 0;
        }
    }
    return true;
}

// Row-relative offsets first (square stencil operators need no per-row base at all), then offsets
// from the first stored column (P, R).  Returns false when the matrix has no such structure.
template <typename TA>
static bool build_patterns(long long n_rows, const int64_t* cp, const int64_t* rv, const TA* nz, int base,
                           bool conjugate, int max_pat, int max_ent, HostPatterns<TA>& out) {
    out = HostPatterns<TA>();
    if (n_rows <= 0) return false;
    if (build_patterns_mode<TA>(n_rows, cp, rv, nz, base, conjugate, true, max_pat, max_ent, out)) return true;
    return build_patterns_mode<TA>(n_rows, cp, rv, nz, base, conjugate, false, max_pat, max_ent, out);
}

// device side ---------------------------------------------------------------------------------
//
// Kernel design (round 1, second iteration; profiles/r01b_*): with one row per thread the kernel is bound by
// the L1 write-back path, not by HBM - every 16-byte dictionary entry is returned to all 32 lanes (512 B per
// warp-load) although the lanes of an interior warp all read the same entry.  The kernel therefore gives every
// thread RPT rows that are `S` rows apart and walks the dictionary entry-outer / row-inner: the entry of the
// thread's first row is loaded once and reused for its other rows whenever they carry the same pattern id (a
// per-lane reload otherwise).  `S` is chosen on the host so that pid[r] == pid[r+S] for most rows: any S works
// for square stencil operators, prolongations repeat with the parity of the node (two grid lines).  Rows are
// still accumulated sequentially in stored order, so results stay bit-identical to the CSR kernels.

template <typename TA>
struct PatDict {
    bool present = false;
    bool rowrel = false;
    int npat = 0, nent = 0;
    int stride = 256;    // S: distance of the rows of one thread
    int nchunk = 1;      // CTAs that share one block of S*RPT rows
    int nthreads = 256;  // threads per CTA (nchunk * nthreads >= S)
    uint16_t* pid = nullptr;
    int* c0 = nullptr;
    int* hdr = nullptr;
    PatEntry<TA>* ent = nullptr;
    std::vector<uint16_t> host_pid;  // kept for the d-folding check at upload
    void release() {
        if (pid) cudaFree(pid);
        if (c0) cudaFree(c0);
        if (hdr) cudaFree(hdr);
        if (ent) cudaFree(ent);
        pid = nullptr;
        c0 = nullptr;
        hdr = nullptr;
        ent = nullptr;
        present = false;
        npat = nent = 0;
        host_pid.clear();
        host_pid.shrink_to_fit();
    }
    // bytes the format really streams per pass, matrix part only (the dictionary itself is cache resident)
    double matrix_bytes(long long n_rows) const { return (double)n_rows * (rowrel ? 2.0 : 6.0); }
};

// Row distance S >= 128 at which the pattern ids repeat (sampled), and the CTA shape that covers S rows.
static inline void choose_pattern_stride(const std::vector<uint16_t>& pid, int& S, int& nchunk, int& nthreads) {
    const long long n = (long long)pid.size();
    S = 256;
    nchunk = 1;
    nthreads = 256;
    if (n < 8192) return;
    const int SMAX = 4160;
    const int nsamp = 4096;
    const long long step = std::max<long long>(1, (n - SMAX) / nsamp);
    auto score = [&](int s) {
        int hit = 0, tot = 0;
        for (long long r = 0; r + s < n && tot < nsamp; r += step, ++tot) hit += (pid[r] == pid[r + s]);
        return tot ? (double)hit / tot : 0.0;
    };
    if (score(256) >= 0.8) return;
    double best = -1.0;
    std::vector<double> sc(SMAX + 1, 0.0);
    for (int s = 128; s <= SMAX; ++s) {
        sc[s] = score(s);
        best = std::max(best, sc[s]);
    }
    if (best < 0.5) return;  // no useful period: any stride is as good as another
    // among the strides close to the best score take the one that wastes the fewest lanes
    double best_cost = 1e30;
    for (int s = 128; s <= SMAX; ++s) {
        if (sc[s] < best - 0.03) continue;
        const int nc = (s + 255) / 256;
        const int nt = (((s + nc - 1) / nc) + 31) / 32 * 32;
        const double cost = (double)nc * nt / s + 1e-5 * s;
        if (cost < best_cost) {
            best_cost = cost;
            S = s;
            nchunk = nc;
            nthreads = nt;
        }
    }
}

template <typename TA>
static void upload_patterns(PatDict<TA>& D, const HostPatterns<TA>& H, long long n_rows) {
    D.release();
    D.rowrel = H.rowrel;
    D.npat = H.npat();
    D.nent = (int)H.delta.size();
    int max_len = 0;
    std::vector<int> hdr(D.npat);
    for (int p = 0; p < D.npat; ++p) {
        const int len = H.pat_off[p + 1] - H.pat_off[p];
        max_len = std::max(max_len, len);
        hdr[p] = H.pat_off[p] | (len << PAT_LEN_SHIFT);
    }
    MGB_CHECK(max_len <= PAT_MAX_LEN && D.nent < (1 << PAT_LEN_SHIFT), "pattern dictionary out of range");
    choose_pattern_stride(H.pid, D.stride, D.nchunk, D.nthreads);
    if (const char* s = std::getenv("MGB200_PAT_STRIDE")) {
        D.stride = std::max(32, std::atoi(s));
        D.nchunk = (D.stride + 255) / 256;
        D.nthreads = (((D.stride + D.nchunk - 1) / D.nchunk) + 31) / 32 * 32;
    }
    MGB_CUDA(cudaMalloc(&D.pid, std::max<size_t>(n_rows, 1) * sizeof(uint16_t)));
    MGB_CUDA(cudaMemcpy(D.pid, H.pid.data(), n_rows * sizeof(uint16_t), cudaMemcpyHostToDevice));
    if (!H.rowrel) {
        MGB_CUDA(cudaMalloc(&D.c0, std::max<size_t>(n_rows, 1) * sizeof(int)));
        MGB_CUDA(cudaMemcpy(D.c0, H.c0.data(), n_rows * sizeof(int), cudaMemcpyHostToDevice));
    }
    MGB_CUDA(cudaMalloc(&D.hdr, std::max(D.npat, 1) * sizeof(int)));
    MGB_CUDA(cudaMemcpy(D.hdr, hdr.data(), D.npat * sizeof(int), cudaMemcpyHostToDevice));
    std::vector<PatEntry<TA>> e((size_t)D.nent + 1);
    std::memset(static_cast<void*>(e.data()), 0, e.size() * sizeof(PatEntry<TA>));
    for (int k = 0; k < D.nent; ++k) {
        e[k].v = H.val[k];
        e[k].delta = H.delta[k];
    }
    MGB_CUDA(cudaMalloc(&D.ent, e.size() * sizeof(PatEntry<TA>)));
    MGB_CUDA(cudaMemcpy(D.ent, e.data(), e.size() * sizeof(PatEntry<TA>), cudaMemcpyHostToDevice));
    D.host_pid = H.pid;
    D.present = true;
}

// ---- loads as volatile inline PTX --------------------------------------------------------------------
// The kernel is latency bound unless every thread keeps many gathers in flight.  nvcc sinks ordinary loads
// next to their first use (one load - one product - next load ...), so the loads are written as volatile
// asm statements, which keep their program order: first all dictionary entries of a chunk, then all gathers,
// then a pin per gathered value that keeps the arithmetic behind the last load.
__device__ __forceinline__ void ld_pred(double& v, const double* p, bool ok) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %2, 0;\n\t@q ld.global.nc.f64 %0, [%1];\n\t}"
                 : "+d"(v) : "l"(p), "r"((int)ok));
}
__device__ __forceinline__ void ld_pred(cplx& v, const cplx* p, bool ok) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %3, 0;\n\t@q ld.global.nc.v2.f64 {%0, %1}, [%2];\n\t}"
                 : "+d"(v.x), "+d"(v.y) : "l"(p), "r"((int)ok));
}
// coherent variant (y of MODE_ADD is read and written by the same kernel)
__device__ __forceinline__ void ld_pred_rw(double& v, const double* p, bool ok) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %2, 0;\n\t@q ld.global.f64 %0, [%1];\n\t}"
                 : "+d"(v) : "l"(p), "r"((int)ok));
}
__device__ __forceinline__ void ld_pred_rw(cplx& v, const cplx* p, bool ok) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %3, 0;\n\t@q ld.global.v2.f64 {%0, %1}, [%2];\n\t}"
                 : "+d"(v.x), "+d"(v.y) : "l"(p), "r"((int)ok));
}
__device__ __forceinline__ void ld_ent_pred(PatEntry<double>& e, const PatEntry<double>* p, bool ok) {
    double t = 0.0;
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %3, 0;\n\t@q ld.global.nc.v2.f64 {%0, %1}, [%2];\n\t}"
                 : "+d"(e.v), "+d"(t) : "l"(p), "r"((int)ok));
    e.delta = __double2loint(t);
}
__device__ __forceinline__ void ld_ent_pred(PatEntry<cplx>& e, const PatEntry<cplx>* p, bool ok) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %4, 0;\n\t@q ld.global.nc.v2.f64 {%0, %1}, [%3];\n\t"
                 "@q ld.global.nc.s32 %2, [%3+16];\n\t}"
                 : "+d"(e.v.x), "+d"(e.v.y), "+r"(e.delta) : "l"(p), "r"((int)ok));
}
__device__ __forceinline__ void pin(double& v) { asm volatile("" : "+d"(v)); }
__device__ __forceinline__ void pin(cplx& v) { asm volatile("" : "+d"(v.x), "+d"(v.y)); }
__device__ __forceinline__ int ld_nc_u16(const uint16_t* p) {
    unsigned short v;
    asm volatile("ld.global.nc.u16 %0, [%1];" : "=h"(v) : "l"(p));
    return (int)v;
}
__device__ __forceinline__ int ld_nc_s32(const int* p) {
    int v;
    asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// acc[j] += sum_k e[k].v * x[base[j] + e[k].delta] for NR rows that share the stencil e[0..L), in stored order.
// Chunks of KC entries: KC dictionary loads, then NR*KC gathers, then the products.
template <typename TA, typename TV>
struct PatChunk {
    static constexpr int KC = (sizeof(TA) > 8 || sizeof(TV) > 8) ? 4 : 8;
};
template <typename TA, typename TV, int NR>
__device__ __forceinline__ void pat_rows(const PatEntry<TA>* __restrict__ e, int L, const int* base,
                                         const TV* __restrict__ x, TV* acc) {
    constexpr int KC = PatChunk<TA, TV>::KC;
    for (int k0 = 0; k0 < L; k0 += KC) {
        PatEntry<TA> ee[KC];
        TV xv[KC][NR];
#pragma unroll
        for (int kk = 0; kk < KC; ++kk) {
            ee[kk].v = VT<TA>::zero();
            ee[kk].delta = 0;
            ld_ent_pred(ee[kk], e + k0 + kk, k0 + kk < L);
        }
#pragma unroll
        for (int kk = 0; kk < KC; ++kk) {
#pragma unroll
            for (int j = 0; j < NR; ++j) {
                xv[kk][j] = VT<TV>::zero();
                ld_pred(xv[kk][j], x + (base[j] + ee[kk].delta), k0 + kk < L);
            }
        }
#pragma unroll
        for (int kk = 0; kk < KC; ++kk) {
#pragma unroll
            for (int j = 0; j < NR; ++j) pin(xv[kk][j]);
        }
#pragma unroll
        for (int kk = 0; kk < KC; ++kk) {
            if (k0 + kk < L) {
#pragma unroll
                for (int j = 0; j < NR; ++j) acc[j] = acc[j] + ee[kk].v * xv[kk][j];
            }
        }
    }
}

// y = op(M x) for one right-hand side; MODE as in csr_kernels.cuh (0 SPMV, 1 ADD, 2 RESID, 3 SWEEP).
// DPAT: the relaxation weights come from the dictionary (dpat[pid]) instead of the vector d.
// Thread t of CTA (c_hi, c_lo) owns rows  c_hi*S*RPT + j*S + c_lo*blockDim + t,  j < RPT.
template <typename TA, typename TV, int MODE, bool ROWREL, bool DPAT, int RPT>
__global__ void __launch_bounds__(256)
pat_kernel(int n_rows, int S, int nchunk, const uint16_t* __restrict__ pid, const int* __restrict__ c0,
           const int* __restrict__ hdr, const PatEntry<TA>* __restrict__ ent, const TV* __restrict__ dpat,
           const TV* __restrict__ x, const TV* __restrict__ b, const TV* __restrict__ d, TV* __restrict__ y) {
    const int c_lo = blockIdx.x % nchunk, c_hi = blockIdx.x / nchunk;
    const int q = c_lo * blockDim.x + threadIdx.x;
    const long long row0 = (long long)c_hi * S * RPT + q;
    if (q >= S || row0 >= n_rows) return;
    int row[RPT], p[RPT], off[RPT], len[RPT], base[RPT];
    bool ok[RPT];
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
        const long long r = row0 + (long long)j * S;
        ok[j] = r < n_rows;
        row[j] = ok[j] ? (int)r : (int)row0;   // rows past the end alias the first one and are never stored
        p[j] = ld_nc_u16(pid + row[j]);
    }
    TV bval[RPT], xval[RPT], dval[RPT], acc[RPT];
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
        acc[j] = VT<TV>::zero();
        bval[j] = VT<TV>::zero();
        xval[j] = VT<TV>::zero();
        dval[j] = VT<TV>::zero();
        if (MODE == 2 || MODE == 3) ld_pred(bval[j], b + row[j], ok[j]);
        if (MODE == 3) ld_pred(xval[j], x + row[j], ok[j]);
        if (MODE == 3 && !DPAT) ld_pred(dval[j], d + row[j], ok[j]);
        if (MODE == 1) ld_pred_rw(xval[j], y + row[j], ok[j]);
        base[j] = ROWREL ? row[j] : ld_nc_s32(c0 + row[j]);
    }
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
        const int h = ld_nc_s32(hdr + p[j]);
        off[j] = h & ((1 << PAT_LEN_SHIFT) - 1);
        len[j] = ok[j] ? (h >> PAT_LEN_SHIFT) : 0;
        if (MODE == 3 && DPAT) ld_pred(dval[j], dpat + p[j], true);
    }
    bool uni = ok[RPT - 1];
#pragma unroll
    for (int j = 1; j < RPT; ++j) uni = uni && (p[j] == p[0]);
    if (uni) {
        // all rows of the thread share one stencil: one dictionary load serves RPT independent gathers
        pat_rows<TA, TV, RPT>(ent + off[0], len[0], base, x, acc);
    } else {
#pragma unroll
        for (int j = 0; j < RPT; ++j) pat_rows<TA, TV, 1>(ent + off[j], len[j], base + j, x, acc + j);
    }
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
        if (!ok[j]) break;
        const int r = row[j];
        if (MODE == 0) {
            y[r] = acc[j];
        } else if (MODE == 1) {
            y[r] = xval[j] + acc[j];
        } else if (MODE == 2) {
            y[r] = bval[j] - acc[j];
        } else {
            const TV res = bval[j] - acc[j];
            y[r] = xval[j] + dval[j] * res;
        }
    }
}

// x = d .* b with d from the dictionary (first sweep of a cycle from x = 0)
template <typename TV>
__global__ void diag_scale_pat_kernel(long long n, const uint16_t* __restrict__ pid, const TV* __restrict__ dpat,
                                      const TV* __restrict__ b, TV* __restrict__ x) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const int p = __ldg(reinterpret_cast<const unsigned short*>(pid) + i);
        x[i] = VT<TV>::zero() + ldg_(dpat + p) * b[i];
    }
}

}  // namespace mgb200
