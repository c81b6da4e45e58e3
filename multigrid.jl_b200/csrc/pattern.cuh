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
    for (long long r : grep) entries += cp[r + 1] - cp[r];
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
template <typename TA>
struct PatDict {
    bool present = false;
    bool rowrel = false;
    int npat = 0, nent = 0;
    uint16_t* pid = nullptr;
    int* c0 = nullptr;
    int* pat_off = nullptr;
    PatEntry<TA>* ent = nullptr;
    std::vector<uint16_t> host_pid;  // kept for the d-folding check at upload
    void release() {
        if (pid) cudaFree(pid);
        if (c0) cudaFree(c0);
        if (pat_off) cudaFree(pat_off);
        if (ent) cudaFree(ent);
        pid = nullptr;
        c0 = nullptr;
        pat_off = nullptr;
        ent = nullptr;
        present = false;
        npat = nent = 0;
        host_pid.clear();
        host_pid.shrink_to_fit();
    }
    // bytes the format really streams per pass, matrix part only (the dictionary itself is cache resident)
    double matrix_bytes(long long n_rows) const { return (double)n_rows * (rowrel ? 2.0 : 6.0); }
};

template <typename TA>
static void upload_patterns(PatDict<TA>& D, const HostPatterns<TA>& H, long long n_rows) {
    D.release();
    D.rowrel = H.rowrel;
    D.npat = H.npat();
    D.nent = (int)H.delta.size();
    MGB_CUDA(cudaMalloc(&D.pid, std::max<size_t>(n_rows, 1) * sizeof(uint16_t)));
    MGB_CUDA(cudaMemcpy(D.pid, H.pid.data(), n_rows * sizeof(uint16_t), cudaMemcpyHostToDevice));
    if (!H.rowrel) {
        MGB_CUDA(cudaMalloc(&D.c0, std::max<size_t>(n_rows, 1) * sizeof(int)));
        MGB_CUDA(cudaMemcpy(D.c0, H.c0.data(), n_rows * sizeof(int), cudaMemcpyHostToDevice));
    }
    MGB_CUDA(cudaMalloc(&D.pat_off, (D.npat + 1) * sizeof(int)));
    MGB_CUDA(cudaMemcpy(D.pat_off, H.pat_off.data(), (D.npat + 1) * sizeof(int), cudaMemcpyHostToDevice));
    std::vector<PatEntry<TA>> e(std::max(D.nent, 1));
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

__device__ __forceinline__ PatEntry<double> ldg_ent(const PatEntry<double>* p) {
    const int4 q = __ldg(reinterpret_cast<const int4*>(p));
    PatEntry<double> e;
    e.v = __hiloint2double(q.y, q.x);
    e.delta = q.z;
    return e;
}
__device__ __forceinline__ PatEntry<cplx> ldg_ent(const PatEntry<cplx>* p) {
    const double2 a = __ldg(reinterpret_cast<const double2*>(p));
    const int4 q = __ldg(reinterpret_cast<const int4*>(p) + 1);
    PatEntry<cplx> e;
    e.v = make_cplx(a.x, a.y);
    e.delta = q.x;
    return e;
}

// y = op(M x) for one right-hand side; MODE as in csr_kernels.cuh (0 SPMV, 1 ADD, 2 RESID, 3 SWEEP).
// DPAT: the relaxation weights come from the dictionary (dpat[pid]) instead of the vector d.
template <typename TA, typename TV, int MODE, bool ROWREL, bool DPAT>
__global__ void __launch_bounds__(256)
pat_kernel(int n_rows, const uint16_t* __restrict__ pid, const int* __restrict__ c0,
           const int* __restrict__ pat_off, const PatEntry<TA>* __restrict__ ent,
           const TV* __restrict__ dpat, const TV* __restrict__ x, const TV* __restrict__ b,
           const TV* __restrict__ d, TV* __restrict__ y) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n_rows) return;
    const int p = __ldg(reinterpret_cast<const unsigned short*>(pid) + row);
    const int base = ROWREL ? row : __ldg(c0 + row);
    const int k0 = __ldg(pat_off + p), k1 = __ldg(pat_off + p + 1);
    TV bval = VT<TV>::zero(), dval = VT<TV>::zero(), xval = VT<TV>::zero();
    if (MODE == 2 || MODE == 3) bval = b[row];
    if (MODE == 3) {
        dval = DPAT ? ldg_(dpat + p) : d[row];
        xval = x[row];
    }
    if (MODE == 1) xval = y[row];
    TV acc = VT<TV>::zero();
#pragma unroll 4
    for (int k = k0; k < k1; ++k) {
        const PatEntry<TA> e = ldg_ent(ent + k);
        acc = acc + e.v * ldg_(x + (base + e.delta));
    }
    if (MODE == 0) {
        y[row] = acc;
    } else if (MODE == 1) {
        y[row] = xval + acc;
    } else if (MODE == 2) {
        y[row] = bval - acc;
    } else {
        const TV r = bval - acc;
        y[row] = xval + dval * r;
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
