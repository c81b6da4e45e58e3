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
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <thread>
#include <unordered_map>
#include <vector>

#include "common.cuh"
#include "ll.cuh"

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
    static_assert(sizeof(TA) % 4 == 0, "value type is made of 4-byte words");
    constexpr int W = sizeof(TA) / 4;
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
            const uint32_t* vb = reinterpret_cast<const uint32_t*>(nz + k0);
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

// Box structure of a row-relative dictionary: every column offset is dz*S2 + dy*S + dx with dx, dy, dz in {-1,0,1}
// (S = line length, S2 = plane length of a lexicographic grid; S2 = 0: no entry leaves the plane, S = 0: none leaves
// the line).  A pattern is then a 27-bit presence mask, bit (dz+1)*9 + (dy+1)*3 + (dx+1), and its entries in stored
// (ascending column) order are its set bits in ascending order.  This is what a line-blocked kernel needs
// (tools/microbench_lines.cu); nothing else about the grid is assumed.
struct BoxInfo {
    bool ok = false;
    int S = 0, S2 = 0;
    std::vector<int> mask;   // per pattern
};
static inline bool box_decompose(long long d, int S, int S2, int& dx, int& dy, int& dz) {
    dz = 0;
    if (S2 > 0) {
        dz = (2 * d > S2) ? 1 : ((2 * d < -(long long)S2) ? -1 : 0);
        d -= (long long)dz * S2;
    }
    dy = 0;
    if (S > 0) {
        dy = (2 * d > S) ? 1 : ((2 * d < -(long long)S) ? -1 : 0);
        d -= (long long)dy * S;
    }
    if (d < -1 || d > 1) return false;
    dx = (int)d;
    return true;
}
template <typename TA>
static BoxInfo detect_box(const HostPatterns<TA>& H, long long n_rows) {
    BoxInfo B;
    if (!H.ok || !H.rowrel || H.delta.empty()) return B;
    std::vector<long long> pos;
    for (int d : H.delta)
        if (d != 0) pos.push_back(d < 0 ? -(long long)d : d);
    std::sort(pos.begin(), pos.end());
    pos.erase(std::unique(pos.begin(), pos.end()), pos.end());
    auto valid = [&](int S, int S2) {
        if (S > 0 && S < 3) return false;
        if (S2 > 0 && (S == 0 || S2 < 3 * S)) return false;
        for (int d : H.delta) {
            int dx, dy, dz;
            if (!box_decompose(d, S, S2, dx, dy, dz)) return false;
            if ((long long)dz * S2 + (long long)dy * S + dx != d) return false;
        }
        return true;
    };
    auto finish = [&](int S, int S2) {
        B.S = S;
        B.S2 = S2;
        B.mask.assign(H.npat(), 0);
        for (int p = 0; p < H.npat(); ++p) {
            int last = -1;
            for (int k = H.pat_off[p]; k < H.pat_off[p + 1]; ++k) {
                int dx, dy, dz;
                box_decompose(H.delta[k], S, S2, dx, dy, dz);
                const int bit = (dz + 1) * 9 + (dy + 1) * 3 + (dx + 1);
                if (bit <= last) return;        // stored order is not (dz,dy,dx) order, or a duplicate offset
                last = bit;
                B.mask[p] |= 1 << bit;
            }
        }
        B.ok = true;
    };
    // candidates: the smallest offset beyond 1 is S-1, S or S+1; the smallest beyond S+1 is S2 + dy*S + dx
    std::vector<long long> big;
    for (long long d : pos)
        if (d > 1) big.push_back(d);
    if (big.empty()) {            // offsets within {-1, 0, 1}: a 1-D stencil
        if (valid(0, 0)) finish(0, 0);
        return B;
    }
    // Several (S, S2) can decompose a sparse offset set (a 7-point stencil on planes of 257 x 257 also reads as lines
    // of 256 with offsets (dy,dx) = (1,1) and (dz,dy,dx) = (1,1,1), and a z-slab of 256 such planes even tiles that
    // way).  Take the simplest reading - the smallest sum of |dz| + |dy| + |dx| over the offsets - and among equals the
    // one that tiles the matrix.
    int best_S = -1, best_S2 = -1, best_score = -1;
    long long best_cost = -1;
    auto cost_of = [&](int S, int S2) {
        std::vector<int> ds(H.delta);
        std::sort(ds.begin(), ds.end());
        ds.erase(std::unique(ds.begin(), ds.end()), ds.end());
        long long c = 0;
        for (int d : ds) {
            int dx, dy, dz;
            box_decompose(d, S, S2, dx, dy, dz);
            c += std::abs(dx) + std::abs(dy) + std::abs(dz);
        }
        return c;
    };
    for (int a = -1; a <= 1; ++a) {
        const long long S = big[0] + a;
        if (S < 3 || S > 0x3fffffff) continue;
        std::vector<long long> bigger;
        for (long long d : big)
            if (d > S + 1) bigger.push_back(d);
        std::vector<long long> cand2;
        if (bigger.empty()) cand2.push_back(0);
        else
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx) cand2.push_back(bigger[0] - dy * S - dx);
        for (long long S2 : cand2) {
            if (S2 < 0 || S2 > 0x3fffffff) continue;
            if (!valid((int)S, (int)S2)) continue;
            // several decompositions can fit a sparse offset set: prefer the one that tiles the matrix
            int score = 0;
            if (S2 > 0 && S2 % S == 0) score += 2;
            if (n_rows % (S2 > 0 ? S2 : S) == 0) score += 1;
            const long long cost = cost_of((int)S, (int)S2);
            if (best_cost < 0 || cost < best_cost || (cost == best_cost && score > best_score)) {
                best_cost = cost;
                best_score = score;
                best_S = (int)S;
                best_S2 = (int)S2;
            }
        }
    }
    if (best_cost >= 0) finish(best_S, best_S2);
    return B;
}

// device side ---------------------------------------------------------------------------------
// Plan of the TMA-staged variant of the kernel (pat_tma_kernel): the column offsets of the whole dictionary are
// clustered into windows [lo, hi]; a CTA that owns a tile of `tile` consecutive rows needs
// x[row0 + lo .. row0 + tile - 1 + hi] of every window - one bulk copy each.
constexpr int TMA_MAX_WIN = 16;
constexpr int TMA_MAX_ENT = 1024;   // dictionary entries / patterns that fit the shared-memory copy
constexpr int TMA_MAX_PAT = 256;
struct TmaWin {
    int lo_even;   // lowest offset of the window rounded down to 16 bytes (2 elements, 4 for Float32)
    int len;       // elements copied (a multiple of the same granularity)
    int sbase;     // first element of the window in the stage buffer
};
struct TmaPlan {
    TmaWin w[TMA_MAX_WIN];
    int nwin, total, centre, tile;
};

template <typename TA>
struct PatDict {
    bool present = false;
    bool rowrel = false;
    int npat = 0, nent = 0;
    // TMA variant (row-relative matrices only)
    bool tma_ok = false;
    TmaPlan plan;
    int* hdr = nullptr;               // entry offset | row length << 20
    PatEntry<TA>* ent_s = nullptr;    // entries whose `delta` is the offset inside the stage buffer
    long long xlo = 0, xhi = 0;       // elements of an input vector that may be copied: [xlo, xhi), both even
    uint16_t* pid = nullptr;
    int* c0 = nullptr;
    int* pat_off = nullptr;
    PatEntry<TA>* ent = nullptr;
    int* rep = nullptr;               // representative row of each pattern (replace_matrix, galerkin.cuh)
    std::vector<uint16_t> host_pid;  // kept for the d-folding check at upload
    // box structure (detect_box): offsets are dz*S2 + dy*S + dx; box_mask[p] = presence bits of pattern p (device)
    bool box_ok = false;
    int S = 0, S2 = 0;
    int* box_mask = nullptr;
    void release() {
        if (box_mask) cudaFree(box_mask);
        box_mask = nullptr;
        box_ok = false;
        S = S2 = 0;
        if (pid) cudaFree(pid);
        if (c0) cudaFree(c0);
        if (pat_off) cudaFree(pat_off);
        if (ent) cudaFree(ent);
        if (rep) cudaFree(rep);
        rep = nullptr;
        if (hdr) cudaFree(hdr);
        if (ent_s) cudaFree(ent_s);
        hdr = nullptr;
        ent_s = nullptr;
        tma_ok = false;
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
struct TmaTile {
    static constexpr int NT = sizeof(TA) > 8 ? 512 : 1024;   // threads per CTA == rows per tile
};
// windows for a tile of `tile` rows; returns false when the offsets need too many windows
template <typename TA>
static bool build_tma_plan(const HostPatterns<TA>& H, int tile, TmaPlan& P, std::vector<int>& soff) {
    if (!H.rowrel || H.delta.empty() || H.npat() > TMA_MAX_PAT || (int)H.delta.size() > TMA_MAX_ENT) return false;
    std::vector<int> ds(H.delta);
    ds.push_back(0);   // the centre (x[row]) is read by the sweep even if the stencil had no diagonal
    std::sort(ds.begin(), ds.end());
    ds.erase(std::unique(ds.begin(), ds.end()), ds.end());
    std::memset(&P, 0, sizeof(P));
    int nw = 0, sb = 0;
    auto close = [&](int lo, int hi) {
        constexpr int AL = sizeof(TA) >= 8 ? 2 : 4;          // elements per 16 bytes (at least 2): tma_align, hierarchy.cuh
        const int lo_even = lo & ~(AL - 1);
        int len = (hi - lo_even) + tile;
        len = (len + AL - 1) & ~(AL - 1);
        P.w[nw].lo_even = lo_even;
        P.w[nw].len = len;
        P.w[nw].sbase = sb;
        sb += len;
        ++nw;
    };
    // Two windows cost 2 * tile + their spans, the merged one tile + spans + gap: merge whenever the gap between
    // neighbouring offsets is below the tile length.  For a lexicographic 3-D stencil the y-neighbour lines then share
    // the copy of the centre line (3 windows instead of 5 for the 7-point level of cfg2, 3 instead of 9 for the 27-point
    // levels): less L2 -> shared-memory traffic and a stage small enough for 2 CTAs per SM on the 27-point levels.
    // MGB200_TMA_GAP restores another threshold (32 = the round-1 plan) for A/B measurements.
    const char* gs = std::getenv("MGB200_TMA_GAP");
    const int gap = gs ? std::atoi(gs) : tile - 1;
    int lo = ds[0], hi = ds[0];
    for (size_t i = 1; i < ds.size(); ++i) {
        if (ds[i] - hi <= gap) {
            hi = ds[i];
            continue;
        }
        if (nw >= TMA_MAX_WIN - 1) return false;
        close(lo, hi);
        lo = hi = ds[i];
    }
    close(lo, hi);
    P.nwin = nw;
    P.total = sb;
    P.tile = tile;
    auto where = [&](int delta) {
        int g = 0;
        while (!(delta >= P.w[g].lo_even && delta <= P.w[g].lo_even + P.w[g].len - tile)) ++g;
        return P.w[g].sbase + (delta - P.w[g].lo_even);
    };
    soff.resize(H.delta.size());
    for (size_t k = 0; k < H.delta.size(); ++k) soff[k] = where(H.delta[k]);
    P.centre = where(0);
    return true;
}

template <typename TA>
static void upload_patterns(PatDict<TA>& D, const HostPatterns<TA>& H, long long n_rows) {
    D.release();
    D.rowrel = H.rowrel;
    D.npat = H.npat();
    D.nent = (int)H.delta.size();
    // 16 ids of slack: the TMA variant copies pid tiles in multiples of 8 ids
    MGB_CUDA(cudaMalloc(&D.pid, (std::max<size_t>(n_rows, 1) + 16) * sizeof(uint16_t)));
    MGB_CUDA(cudaMemset(D.pid, 0, (std::max<size_t>(n_rows, 1) + 16) * sizeof(uint16_t)));
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
    if ((int)H.rep_row.size() == D.npat) {
        std::vector<int> rr(H.rep_row.begin(), H.rep_row.end());
        MGB_CUDA(cudaMalloc(&D.rep, std::max(D.npat, 1) * sizeof(int)));
        MGB_CUDA(cudaMemcpy(D.rep, rr.data(), D.npat * sizeof(int), cudaMemcpyHostToDevice));
    }
    D.host_pid = H.pid;
    D.present = true;
    {
        const BoxInfo B = detect_box<TA>(H, n_rows);
        if (B.ok && B.S >= 3) {   // kept for the box-stencil kernel (box.cuh) and mgb200_host_detect_box
            MGB_CUDA(cudaMalloc(&D.box_mask, D.npat * sizeof(int)));
            MGB_CUDA(cudaMemcpy(D.box_mask, B.mask.data(), D.npat * sizeof(int), cudaMemcpyHostToDevice));
            D.box_ok = true;
            D.S = B.S;
            D.S2 = B.S2;
        }
    }
    // ---- TMA variant ----
    std::vector<int> soff;
    int max_len = 0;
    for (int p = 0; p < D.npat; ++p) max_len = std::max(max_len, H.pat_off[p + 1] - H.pat_off[p]);
    if (max_len < (1 << 11) && build_tma_plan<TA>(H, TmaTile<TA>::NT, D.plan, soff)) {
        std::vector<int> hdr(D.npat);
        for (int p = 0; p < D.npat; ++p) hdr[p] = H.pat_off[p] | ((H.pat_off[p + 1] - H.pat_off[p]) << 20);
        MGB_CUDA(cudaMalloc(&D.hdr, D.npat * sizeof(int)));
        MGB_CUDA(cudaMemcpy(D.hdr, hdr.data(), D.npat * sizeof(int), cudaMemcpyHostToDevice));
        for (int k = 0; k < D.nent; ++k) e[k].delta = soff[k];
        MGB_CUDA(cudaMalloc(&D.ent_s, e.size() * sizeof(PatEntry<TA>)));
        MGB_CUDA(cudaMemcpy(D.ent_s, e.data(), e.size() * sizeof(PatEntry<TA>), cudaMemcpyHostToDevice));
        D.tma_ok = true;
    }
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

__device__ __forceinline__ PatEntry<float> ldg_ent(const PatEntry<float>* p) {
    const int2 q = __ldg(reinterpret_cast<const int2*>(p));
    PatEntry<float> e;
    e.v = __int_as_float(q.x);
    e.delta = q.y;
    return e;
}
__device__ __forceinline__ PatEntry<cplxf> ldg_ent(const PatEntry<cplxf>* p) {
    const int4 q = __ldg(reinterpret_cast<const int4*>(p));
    PatEntry<cplxf> e;
    e.v = make_cplxf(__int_as_float(q.x), __int_as_float(q.y));
    e.delta = q.z;
    return e;
}

// y = op(M x) for one right-hand side; MODE as in csr_kernels.cuh (0 SPMV, 1 ADD, 2 RESID, 3 SWEEP).
// DPAT: the relaxation weights come from the dictionary (dpat[pid]) instead of the vector d.
// The launch covers two row ranges, [rA, rA + nA) and then [rB, rB + nB) (a whole matrix is (0, n, 0, 0)); the
// split form serves the rows next to the slab ends after a halo exchange that ran beside the interior rows.
template <typename TA, typename TV, int MODE, bool ROWREL, bool DPAT>
__global__ void __launch_bounds__(256)
pat_kernel(const __grid_constant__ PutPlan pp, int rA, int nA, int rB, int nB, const uint16_t* __restrict__ pid,
           const int* __restrict__ c0,
           const int* __restrict__ pat_off, const PatEntry<TA>* __restrict__ ent,
           const TV* __restrict__ dpat, const TV* __restrict__ x, const TV* __restrict__ b,
           const TV* __restrict__ d, TV* __restrict__ y) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nA + nB) return;
    const int row = idx < nA ? rA + idx : rB + (idx - nA);
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
    TV out;
    if (MODE == 0) {
        out = acc;
    } else if (MODE == 1) {
        out = xval + acc;
    } else if (MODE == 2) {
        out = bval - acc;
    } else {
        const TV r = bval - acc;
        out = xval + dval * r;
    }
    y[row] = out;
    if (pp.on) ll_put_edge<TV>(pp, row, out);
}

// ---- host / device helpers shared with box.cuh and with the CPU replays (mgb200_host_pattern_apply) -----------------------
template <typename T>
__host__ __device__ __forceinline__ T ld_ro(const T* p) {
#ifdef __CUDA_ARCH__
    return ldg_(p);
#else
    return *p;
#endif
}
__host__ __device__ __forceinline__ int ld_pid(const uint16_t* p) {
#ifdef __CUDA_ARCH__
    return __ldg(reinterpret_cast<const unsigned short*>(p));
#else
    return *p;
#endif
}
template <int MODE, typename TV>
__host__ __device__ __forceinline__ TV pat_epilogue(TV acc, TV xval, TV bval, TV dval) {
    if (MODE == 0) return acc;
    if (MODE == 2) return bval - acc;
    const TV r = bval - acc;
    return xval + dval * r;
}
// View of the vectors for the one-row dictionary walk below: xs[dz+1][row] == x[dz*S2 + row]; b, d, pid by row.
template <typename TV>
struct LinesView {
    const TV* xs[3];
    const TV* b;
    const TV* d;
    const uint16_t* pid;
};
// one row, dictionary walked entry by entry (what pat_kernel<..., ROWREL = true> computes), x through the view
template <typename TA, typename TV, int MODE, bool DPAT>
__host__ __device__ inline TV pat_row_walk(long long S2, const LinesView<TV>& V, long long rel, const int* pat_off,
                                           const PatEntry<TA>* ent, const TV* dpat) {
    const int p = ld_pid(V.pid + rel);
    const int k0 = ld_ro(pat_off + p), k1 = ld_ro(pat_off + p + 1);
    TV acc = VT<TV>::zero();
    for (int k = k0; k < k1; ++k) {
        const long long dl = ent[k].delta;
        const int dz = (S2 > 0) ? ((2 * dl > S2) ? 1 : ((2 * dl < -S2) ? -1 : 0)) : 0;
        const TV* xb = dz > 0 ? V.xs[2] : (dz < 0 ? V.xs[0] : V.xs[1]);          // no dynamic index into the view
        acc = acc + ent[k].v * ld_ro(xb + (rel + (dl - dz * S2)));
    }
    TV bval = VT<TV>::zero(), dval = VT<TV>::zero(), xval = VT<TV>::zero();
    if (MODE == 2 || MODE == 3) bval = V.b[rel];
    if (MODE == 3) {
        dval = DPAT ? ld_ro(dpat + p) : V.d[rel];
        xval = V.xs[1][rel];
    }
    return pat_epilogue<MODE, TV>(acc, xval, bval, dval);
}
template <typename TV>
__host__ __device__ __forceinline__ LinesView<TV> lines_global_view(long long S2, const uint16_t* pid, const TV* x, const TV* b,
                                                                    const TV* d) {
    LinesView<TV> V;
    V.xs[0] = x - S2;      // pointer arithmetic only: loads go where entries point
    V.xs[1] = x;
    V.xs[2] = x + S2;
    V.b = b;
    V.d = d;
    V.pid = pid;
    return V;
}
// ---- TMA-staged variant ----------------------------------------------------------------------------------
// Persistent CTAs of NT threads walk tiles of NT consecutive rows with a two-stage pipeline: while the CTA computes
// tile i, the bulk copies of tile i+1 (the x windows, the b / d tiles and the pattern ids, issued by one thread,
// cp.async.bulk + mbarrier) are in flight.  The dictionary is copied to shared memory once per CTA.  One thread
// still owns one row and accumulates in stored order: bit-identical to pat_kernel.  (tools/microbench_pat.cu:
// 105 us against 140 us for the 7-point sweep at 257^3.)
// The per-tile __syncthreads is the top stall in ncu (profiles/r01e_ncu_full_dictionary_kernels_summary.txt), but
// releasing the stages per warp through `empty` mbarriers instead (each warp arrives, only the issuing thread
// waits) measured SLOWER: level-1 sweep 119 -> 125 us, level-2 sweep 46 -> 54 us, cycle 1.033 -> 1.090 ms.
template <typename TA, typename TV, int MODE, bool DPAT, int NT>
__global__ void __launch_bounds__(NT)
pat_tma_kernel(const __grid_constant__ TmaPlan P, const __grid_constant__ PutPlan pp, int n_rows, int tile0, int ntiles, long long xlo, long long xhi, int npat,
               int nent, const uint16_t* __restrict__ pid, const int* __restrict__ hdr,
               const PatEntry<TA>* __restrict__ ent_s, const TV* __restrict__ dpat, const TV* __restrict__ x,
               const TV* __restrict__ b, const TV* __restrict__ d, TV* __restrict__ y) {
    constexpr int STAGES = 2;
    constexpr bool NEED_B = (MODE == 2 || MODE == 3);
    constexpr bool NEED_D = (MODE == 3 && !DPAT);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
    PatEntry<TA>* se = reinterpret_cast<PatEntry<TA>*>(smem_raw + 64);
    TV* sdp = reinterpret_cast<TV*>(se + nent);
    int* sh = reinterpret_cast<int*>(sdp + npat);
    unsigned char* stage0 = smem_raw + ((64 + (size_t)nent * sizeof(PatEntry<TA>) + (size_t)npat * (sizeof(TV) + 4) + 127) / 128) * 128;
    const int elems = P.total + (NEED_B ? NT : 0) + (NEED_D ? NT : 0);
    const size_t stage_bytes = (((size_t)elems * sizeof(TV) + (size_t)NT * 2) + 127) / 128 * 128;
    const int t = threadIdx.x;
    if (t == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(full + s, 1);
        fence_mbar_init();
    }
    for (int i = t; i < nent; i += NT) se[i] = ent_s[i];
    for (int i = t; i < npat; i += NT) {
        sh[i] = hdr[i];
        if (MODE == 3 && DPAT) sdp[i] = dpat[i];
    }
    __syncthreads();
    auto issue = [&](int tile, int s) {
        TV* sx = reinterpret_cast<TV*>(stage0 + (size_t)s * stage_bytes);
        TV* sb = sx + P.total;
        TV* sd = sb + (NEED_B ? NT : 0);
        uint16_t* sp = reinterpret_cast<uint16_t*>(sx + elems);
        const long long row0 = (long long)tile * NT;
        constexpr int AL = sizeof(TV) >= 8 ? 2 : 4;
        const long long rend = min(row0 + NT, (long long)((n_rows + AL - 1) & ~(AL - 1)));   // vectors carry 4 elements of slack
        const long long pend = min(row0 + NT, (long long)((n_rows + 7) & ~7));
        uint32_t bytes = (uint32_t)(pend - row0) * 2u;
        for (int g = 0; g < P.nwin; ++g) {
            const long long s0 = row0 + P.w[g].lo_even, a = max(s0, xlo), e = min(s0 + P.w[g].len, xhi);
            if (e > a) bytes += (uint32_t)(e - a) * (uint32_t)sizeof(TV);
        }
        if (NEED_B) bytes += (uint32_t)(rend - row0) * (uint32_t)sizeof(TV);
        if (NEED_D) bytes += (uint32_t)(rend - row0) * (uint32_t)sizeof(TV);
        mbar_expect_tx(full + s, bytes);
        for (int g = 0; g < P.nwin; ++g) {
            const long long s0 = row0 + P.w[g].lo_even, a = max(s0, xlo), e = min(s0 + P.w[g].len, xhi);
            if (e > a) bulk_g2s(sx + P.w[g].sbase + (a - s0), x + a, (uint32_t)(e - a) * (uint32_t)sizeof(TV), full + s);
        }
        if (NEED_B) bulk_g2s(sb, b + row0, (uint32_t)(rend - row0) * (uint32_t)sizeof(TV), full + s);
        if (NEED_D) bulk_g2s(sd, d + row0, (uint32_t)(rend - row0) * (uint32_t)sizeof(TV), full + s);
        bulk_g2s(sp, pid + row0, (uint32_t)(pend - row0) * 2u, full + s);
    };
    // tiles [tile0, ntiles) of the matrix: the whole matrix, or the interior tiles of a split launch
    if (t == 0 && tile0 + (int)blockIdx.x < ntiles) issue(tile0 + blockIdx.x, 0);
    int i = 0;
    for (int tile = tile0 + blockIdx.x; tile < ntiles; tile += gridDim.x, ++i) {
        const int s = i & 1;
        if (t == 0) {
            const int nt = tile + gridDim.x;
            if (nt < ntiles) issue(nt, s ^ 1);      // stage s^1 was released by the barrier that ended iteration i-1
        }
        mbar_wait(full + s, (i >> 1) & 1);
        const TV* sx = reinterpret_cast<const TV*>(stage0 + (size_t)s * stage_bytes);
        const TV* sb = sx + P.total;
        const TV* sd = sb + (NEED_B ? NT : 0);
        const uint16_t* sp = reinterpret_cast<const uint16_t*>(sx + elems);
        const long long row = (long long)tile * NT + t;
        if (row < n_rows) {
            const int p = sp[t];
            const int h = sh[p], k0 = h & 0xFFFFF, k1 = k0 + (h >> 20);
            const TV* sxt = sx + t;
            TV acc = VT<TV>::zero();
#pragma unroll 4
            for (int k = k0; k < k1; ++k) {
                const PatEntry<TA> e = se[k];
                acc = acc + e.v * sxt[e.delta];
            }
            TV out;
            if (MODE == 0) {
                out = acc;
            } else if (MODE == 2) {
                out = sb[t] - acc;
            } else {
                const TV dval = DPAT ? sdp[p] : sd[t];
                const TV res = sb[t] - acc;
                out = sxt[P.centre] + dval * res;
            }
            y[row] = out;
            if (pp.on) ll_put_edge<TV>(pp, row, out);
        }
        __syncthreads();
    }
}

// x = d .* b with d from the dictionary (first sweep of a cycle from x = 0)
template <typename TV>
__global__ void diag_scale_pat_kernel(const __grid_constant__ PutPlan pp, long long n, const uint16_t* __restrict__ pid,
                                      const TV* __restrict__ dpat, const TV* __restrict__ b, TV* __restrict__ x) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const int p = __ldg(reinterpret_cast<const unsigned short*>(pid) + i);
        const TV out = VT<TV>::zero() + ldg_(dpat + p) * b[i];
        x[i] = out;
        if (pp.on) ll_put_edge<TV>(pp, i, out);
    }
}

}  // namespace mgb200
