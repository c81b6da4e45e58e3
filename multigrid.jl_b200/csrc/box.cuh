// Box-stencil kernel: the fused sweep / residual / product of a square operator whose stencil dictionary has box
// structure on a lexicographic 3-D grid (detect_box, pattern.cuh: every column offset is dz*S2 + dy*S + dx with
// dx, dy, dz in {-1, 0, 1}).  The geometric hierarchies of the reference (MGsetup.jl:94-111: 7-point fine operators,
// 27-point Galerkin products, or rediscretised 7-point operators on every level) are exactly that.
//
// Why another kernel: the dictionary walk of pat_tma_kernel (pattern.cuh) spends ~145 warp instructions per 32 rows of
// the 7-point level - header, loop, entry loads with their offsets - and is bound by issue slots and shared-memory
// wavefronts, not by DRAM (profiles/r01i_ncu_full_dictionary_kernels_summary.txt).  Here the stencil SHAPE is a
// compile-time constant (the 7-point star or the full 27-point box), so the product chain is fully unrolled with
// immediate offsets, and the dictionary becomes a DENSE coefficient table coef[k][pattern] with zeros where a
// boundary pattern has no entry.  Adding 0 * x[..] leaves a sum unchanged bit for bit (the skipped neighbour of a
// boundary row is a finite element of the staged window), so one thread still accumulates one row in stored
// (dz, dy, dx) order and the results are bit-identical to pat_kernel and to the CPU oracle.
//  * Warps whose rows all carry the dominant (interior) pattern take the coefficients from the kernel parameter, i.e.
//    straight from the constant bank as instruction operands: no coefficient loads at all.
//  * A thread owns RZ rows that are one plane apart (rows r, r + S2, ...): the x values of a plane are loaded once
//    for the (up to) three rows that multiply them, and a tile needs RZ + 2 plane windows for RZ planes of rows
//    instead of 3 for 1.  Planes are visited in ascending order, which IS stored order for every row.
// Staging is as in pat_tma_kernel: one bulk copy (cp.async.bulk + mbarrier) per plane window, plus the b / d / pattern
// id tiles; persistent CTAs, STAGES = 2 (copies of tile i+1 fly while tile i is computed) or 1 (several CTAs per SM
// cover each other's copies).  The copy list of every tile is planned ON THE HOST once per (matrix, variant) and kept
// in device memory as one record per tile (BoxRec: the copies + a descriptor of the tile's geometry): the issuing warp
// loads one 16-byte copy entry per lane and fires - planning a tile in the kernel (alignment of odd plane lengths,
// clamping at the ends of the vector) put ~200 dependent instructions in front of every tile's copies and made
// the issuing warp the critical path (profiles/r02_box_variants.log).
// The per-thread function and the tile plan are __host__ __device__: mgb200_host_box_apply replays a launch on the CPU
// against host buffers filled exactly where the bulk copies would fill shared memory (tests/test_patterns.py).
#pragma once
#include "pattern.cuh"

namespace mgb200 {

constexpr int BOX_MAX_RZ = 4;
constexpr int BOX_MAX_PAT = 64;      // patterns of a dense coefficient table (27 boundary classes on a box grid)

// stencil shapes: 7 = star (centre, +-1 in each direction), 27 = full box.  Entry index k of (dz,dy,dx) in stored order.
template <int SHAPE>
__host__ __device__ constexpr bool box_has(int dz, int dy, int dx) {
    return SHAPE == 27 ? true : ((dz != 0) + (dy != 0) + (dx != 0) <= 1);
}
template <int SHAPE>
__host__ __device__ constexpr int box_k(int dz, int dy, int dx) {
    if (SHAPE == 27) return (dz + 1) * 9 + (dy + 1) * 3 + (dx + 1);
    // star, ascending column order: (-1,0,0) (0,-1,0) (0,0,-1) (0,0,0) (0,0,1) (0,1,0) (1,0,0)
    return dz < 0 ? 0 : (dz > 0 ? 6 : (dy < 0 ? 1 : (dy > 0 ? 5 : 3 + dx)));
}
constexpr int BOX_STAR_MASK = (1 << 4) | (1 << 10) | (1 << 12) | (1 << 13) | (1 << 14) | (1 << 16) | (1 << 22);

// coefficients of the dominant pattern: a kernel parameter, so the fast path reads them from the constant bank
template <typename TV>
struct BoxCoef {
    TV c[27];
    TV d0;       // folded relaxation weight of the dominant pattern (DPAT)
};

// (32-bit on purpose: the issuing warp plans every tile with these; the launcher checks that n_rows + 2 S2 fits)
struct BoxPlan {
    int n_rows;
    int S, S2;              // line / plane length of the grid
    int plane;              // rows per plane of the tiling: S2, or n_rows when a thread owns one row (RZ == 1)
    int nplanes;            // planes of the tiling: n_rows / S2, or 1
    int xlo, xhi;           // elements of an input vector that may be copied: [xlo, xhi), multiples of the copy granule
    int ntiles;
    int nchunk;             // chunks of NB base rows per plane
    int npat, NP, p0;       // patterns, leading dimension of the coefficient table, dominant pattern
    int wlo[BOX_MAX_RZ + 2], whi[BOX_MAX_RZ + 2];   // window w holds offsets [wlo, NB - 1 + whi] around its centre rows
    int sbase[BOX_MAX_RZ + 2];                      // first element of window w in a stage
    int xtotal;             // elements of all windows of a stage
    int pbase[BOX_MAX_RZ + 2];                      // first pattern id of the id window w (MODE 4 only)
    int ptotal;             // pattern ids of all id windows
    int first_ghost_tile;   // records from this one on read ghost rows of the input vector (they are sorted last)
};
// A consumer that runs beside the halo exchange of its input vector (row-partitioned levels): tiles that read ghost rows
// wait inside the kernel until the exchange kernel has published exchange number *consumed + 1 (p2p.cuh).
struct BoxWait {
    const unsigned long long* epoch;    // null: no waiting
    unsigned long long* consumed;
    unsigned* ticket;
};
static inline BoxWait no_wait() {
    BoxWait w;
    w.epoch = nullptr;
    w.consumed = nullptr;
    w.ticket = nullptr;
    return w;
}
// MODE 4 of the box kernel: the first TWO sweeps of a relaxation that starts from x = 0 (MGcycle.jl:128-135 with the
// reference's `x .= 0` start), x1 = 0 + d.*b, x2 = x1 + d.*(b - A x1), in one pass: the staged windows hold b and the
// pattern ids of the neighbours, and x1 of a neighbour is recomputed where it is needed.  Needs d folded into the
// dictionary (d = dtab[pattern id]).
constexpr int MODE_SWEEP2_FROM_ZERO = 4;

template <typename TV>
__host__ __device__ constexpr int box_al() { return sizeof(TV) >= 8 ? 2 : 4; }   // elements per 16 bytes (tma_align)

// host: windows of a tile for SHAPE / RZ / NB
template <typename TV>
static inline void box_make_plan(BoxPlan& P, int shape, int RZ, int NB, long long n_rows, long long S, long long S2,
                                 long long xlo, long long xhi, int npat, int p0) {
    constexpr int AL = box_al<TV>();
    std::memset(&P, 0, sizeof(P));
    P.n_rows = (int)n_rows;
    P.S = (int)S;
    P.S2 = (int)S2;
    P.plane = (int)(RZ == 1 ? n_rows : S2);
    P.xlo = (int)xlo;
    P.xhi = (int)xhi;
    P.nchunk = (P.plane + NB - 1) / NB;
    P.nplanes = (int)((n_rows + P.plane - 1) / P.plane);
    const long long nt = (long long)((P.nplanes + RZ - 1) / RZ) * P.nchunk;
    P.ntiles = nt < (1LL << 31) ? (int)nt : 0;
    P.npat = npat;
    P.NP = (npat + 1) & ~1;
    P.p0 = p0;
    int sb = 0;
    for (int w = 0; w < RZ + 2; ++w) {
        const bool inner = (w >= 1 && w <= RZ);
        const int span = (shape == 27 || inner) ? (int)(S + 1) : 0;
        P.wlo[w] = -span;
        P.whi[w] = span;
        P.sbase[w] = sb;
        sb += (NB + 2 * span + 2 * AL + AL - 1) / AL * AL;
    }
    P.xtotal = sb;
    int pb = 0;
    for (int w = 0; w < RZ + 2; ++w) {
        P.pbase[w] = pb;
        pb += (NB + P.whi[w] - P.wlo[w] + 16 + 7) / 8 * 8;
    }
    P.ptotal = pb;
}
template <typename TV>
__host__ __device__ inline size_t box_stage_bytes(const BoxPlan& P, int RZ, int NB, bool need_b, bool need_d, bool need_pw = false) {
    constexpr int AL = box_al<TV>();
    const size_t bcap = NB + 2 * AL, pcap = NB + 16;
    const size_t bytes = ((size_t)P.xtotal + (need_b ? RZ * bcap : 0) + (need_d ? RZ * bcap : 0)) * sizeof(TV) + RZ * pcap * 2 +
                         (need_pw ? (size_t)P.ptotal * 2 : 0);
    // [descriptor][x windows][pattern ids][b][d], or in MODE 4 [descriptor][b windows][pattern ids][id windows]
    return 128 /* BOX_DESC_BYTES */ + (bytes + 127) / 128 * 128;
}
template <typename TV>
__host__ __device__ inline size_t box_head_bytes(const BoxPlan& P, int nk) {
    return ((64 + ((size_t)nk * P.NP + P.NP) * sizeof(TV)) + 127) / 128 * 128;
}

// One copy of a tile: `bytes` from global element `src` (of array `what`: 0 x, 1 b, 2 d, 3 pid, 4 the tile's own
// descriptor) to byte offset `dst` of the stage; bytes == 0: nothing to copy.
struct __align__(16) BoxCopy {
    int src;
    unsigned dst, bytes;
    int what;
};
__host__ __device__ inline int box_floor(int a, int al) { return a & ~(al - 1); }
__host__ __device__ inline int box_ceil(int a, int al) { return (a + al - 1) & ~(al - 1); }
// copies of a tile: the windows 0 .. RZ+1, then per row-plane j the pattern ids, the b tile and the d tile, then the
// pattern-id windows (what = 5, MODE 4 only)
__host__ __device__ constexpr int box_ncopies(int RZ) { return RZ + 2 + 3 * RZ + RZ + 2; }
// Stage layout: [descriptor, 128 bytes][x windows][pattern ids][b tiles][d tiles] (b, d only in the modes that read them).
constexpr int BOX_DESC_BYTES = 128;
struct BoxDesc {           // what every thread needs of a tile's geometry: element indices of thread 0's elements
    int r0, nb, nrp, pad_;
    int xoff[BOX_MAX_RZ + 2];   // centre element of window w, from the start of the x region
    int boff[BOX_MAX_RZ];       // row of row-plane j, from the start of the b (or d) region
    int poff[BOX_MAX_RZ];       // same for the pattern ids
    int pwoff[BOX_MAX_RZ + 2];  // centre element of id window w, from the start of the id-window region (MODE 4)
};
static_assert(sizeof(BoxDesc) <= BOX_DESC_BYTES, "descriptor does not fit its slot");
// One record per tile in device memory: the descriptor (copied into the stage like the data) and the copy list.
__host__ __device__ constexpr int box_rec_bytes(int RZ) { return BOX_DESC_BYTES + box_ncopies(RZ) * (int)sizeof(BoxCopy); }

// geometry of tile (plane group g, chunk c): first row, base rows, row-planes that exist
inline void box_tile_rows(const BoxPlan& P, int RZ, int NB, int g, int c, int& r0, int& nb, int& nrp) {
    const int c0 = c * NB;
    r0 = g * RZ * P.plane + c0;
    nb = P.plane - c0 < NB ? P.plane - c0 : NB;
    const int left = P.nplanes - g * RZ;               // RZ > 1: n_rows is a multiple of S2 (launcher)
    nrp = left < RZ ? left : RZ;
}
// host: the record of one tile (rec: box_rec_bytes(RZ) bytes)
template <typename TV>
inline void box_plan_tile(const BoxPlan& P, int RZ, int NB, int tile, unsigned char* rec) {
    constexpr int AL = box_al<TV>();
    std::memset(rec, 0, box_rec_bytes(RZ));
    BoxDesc* D = reinterpret_cast<BoxDesc*>(rec);
    BoxCopy* cp = reinterpret_cast<BoxCopy*>(rec + BOX_DESC_BYTES);
    int r0, nb, nrp;
    box_tile_rows(P, RZ, NB, tile / P.nchunk, tile % P.nchunk, r0, nb, nrp);
    D->r0 = r0;
    D->nb = nb;
    D->nrp = nrp;
    const int bcap = NB + 2 * AL, pcap = NB + 16;
    const size_t x_bytes = BOX_DESC_BYTES;
    const size_t p_bytes = x_bytes + (size_t)P.xtotal * sizeof(TV);
    const size_t b_bytes = p_bytes + (size_t)RZ * pcap * 2;
    const size_t d_bytes = b_bytes + (size_t)RZ * bcap * sizeof(TV);
    int n = 0;
    for (int w = 0; w < RZ + 2; ++w, ++n) {
        BoxCopy& C = cp[n];
        C.what = 0;
        const int centre = r0 + (w - 1) * P.S2;
        const int a0 = box_floor(centre + P.wlo[w], AL);
        D->xoff[w] = P.sbase[w] + (centre - a0);
        if (w - 2 >= nrp) continue;            // none of the window's (up to three) row-planes exists
        int a = a0, e = box_ceil(centre + nb + P.whi[w], AL);
        if (a < P.xlo) a = P.xlo;
        if (e > P.xhi) e = P.xhi;
        if (e <= a) continue;
        C.src = a;
        C.dst = (unsigned)(x_bytes + (size_t)(P.sbase[w] + (a - a0)) * sizeof(TV));
        C.bytes = (unsigned)((e - a) * sizeof(TV));
    }
    const int nal = box_ceil(P.n_rows, AL), n8 = box_ceil(P.n_rows, 8);
    for (int j = 0; j < RZ; ++j) {
        const int r = r0 + j * P.S2;
        const int a = box_floor(r, AL), a8 = box_floor(r, 8);
        D->boff[j] = j * bcap + (r - a);
        D->poff[j] = j * pcap + (r - a8);
        BoxCopy& Cp = cp[n++];
        BoxCopy& Cb = cp[n++];
        BoxCopy& Cd = cp[n++];
        Cp.what = 3;
        Cb.what = 1;
        Cd.what = 2;
        if (j >= nrp) continue;
        int e = box_ceil(r + nb, AL), e8 = box_ceil(r + nb, 8);
        if (e > nal) e = nal;
        if (e8 > n8) e8 = n8;
        Cp.src = a8;
        Cp.dst = (unsigned)(p_bytes + (size_t)j * pcap * 2);
        Cp.bytes = (unsigned)((e8 - a8) * 2);
        Cb.src = Cd.src = a;
        Cb.dst = (unsigned)(b_bytes + (size_t)j * bcap * sizeof(TV));
        Cd.dst = (unsigned)(d_bytes + (size_t)j * bcap * sizeof(TV));
        Cb.bytes = Cd.bytes = (unsigned)((e - a) * sizeof(TV));
    }
    // pattern-id windows (MODE 4: no b / d tiles in the stage, the id windows follow the id tiles)
    const size_t pw_bytes = b_bytes;
    for (int w = 0; w < RZ + 2; ++w, ++n) {
        BoxCopy& C = cp[n];
        C.what = 5;
        const int centre = r0 + (w - 1) * P.S2;
        const int a0 = box_floor(centre + P.wlo[w], 8);
        D->pwoff[w] = P.pbase[w] + (centre - a0);
        if (w - 2 >= nrp) continue;
        int a = a0, e = box_ceil(centre + nb + P.whi[w], 8);
        if (a < 0) a = 0;
        if (e > n8) e = n8;
        if (e <= a) continue;
        C.src = a;
        C.dst = (unsigned)(pw_bytes + (size_t)(P.pbase[w] + (a - a0)) * 2);
        C.bytes = (unsigned)((e - a) * 2);
    }
}

// One thread: RZ rows one plane apart.  xc[w]: the thread's centre element of window w; coefficients from C0 (FAST: every
// row of the warp carries the dominant pattern) or from the dense table ctab[k * NP + pattern].
template <typename TV, int SHAPE, int MODE, bool DPAT, int RZ, bool FAST>
__host__ __device__ __forceinline__ void box_thread(const BoxCoef<TV>& C0, const TV* ctab, const TV* dtab, int NP, int S,
                                                    const TV* const* xc, const TV* const* bp, const TV* const* dp,
                                                    const int* pat, TV* out, const uint16_t* const* pw = nullptr) {
    TV acc[RZ], xcen[RZ], bcen[RZ];
#pragma unroll
    for (int j = 0; j < RZ; ++j) {
        acc[j] = VT<TV>::zero();
        xcen[j] = VT<TV>::zero();
        bcen[j] = VT<TV>::zero();
    }
#pragma unroll
    for (int w = 0; w < RZ + 2; ++w) {
        const bool inner = (w >= 1 && w <= RZ);
        const TV* q = xc[w];
        TV X[3][3];
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
                const bool need = SHAPE == 27 || (dy == 0 && dx == 0) || (inner && (dy == 0 || dx == 0));
                if (need) {
                    if (MODE == 4) {      // the window holds b: x1 = 0 + d .* b, exactly as diag_scale_pat_kernel computes it
                        const TV bv = q[dy * S + dx];
                        if (inner && dy == 0 && dx == 0) bcen[inner ? w - 1 : 0] = bv;
                        X[dy + 1][dx + 1] = VT<TV>::zero() + dtab[pw[w][dy * S + dx]] * bv;
                    } else {
                        X[dy + 1][dx + 1] = q[dy * S + dx];
                    }
                }
            }
#pragma unroll
        for (int j = 0; j < RZ; ++j) {
            const int dz = w - 1 - j;
            if (dz < -1 || dz > 1) continue;
            if (dz == 0) xcen[j] = X[1][1];
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
                for (int dx = -1; dx <= 1; ++dx) {
                    if (!box_has<SHAPE>(dz, dy, dx)) continue;
                    const int k = box_k<SHAPE>(dz, dy, dx);
                    const TV cf = FAST ? C0.c[k] : ctab[k * NP + pat[j]];
                    acc[j] = acc[j] + cf * X[dy + 1][dx + 1];
                }
        }
    }
#pragma unroll
    for (int j = 0; j < RZ; ++j) {
        TV bval = VT<TV>::zero(), dval = VT<TV>::zero();
        if (MODE == 2 || MODE == 3) bval = *bp[j];
        if (MODE == 4) bval = bcen[j];
        if (MODE == 3) dval = DPAT ? (FAST ? C0.d0 : dtab[pat[j]]) : *dp[j];
        if (MODE == 4) dval = FAST ? C0.d0 : dtab[pat[j]];
        out[j] = pat_epilogue<(MODE == 4 ? 3 : MODE), TV>(acc[j], xcen[j], bval, dval);
    }
}

template <typename TV, int SHAPE, int MODE, bool DPAT, int RZ, int NB, int STAGES>
__global__ void __launch_bounds__(NB)
box_kernel(const __grid_constant__ BoxPlan P, const __grid_constant__ BoxCoef<TV> C0, const __grid_constant__ PutPlan pp,
           const __grid_constant__ BoxWait bw, const unsigned char* __restrict__ recs, const uint16_t* __restrict__ pid,
           const TV* __restrict__ ctab_g,
           const TV* __restrict__ dtab_g, const TV* __restrict__ x, const TV* __restrict__ b, const TV* __restrict__ d,
           TV* __restrict__ y) {
    constexpr bool NEED_B = (MODE == 2 || MODE == 3);
    constexpr bool NEED_D = (MODE == 3 && !DPAT);
    constexpr bool NEED_PW = (MODE == 4);
    constexpr int NK = SHAPE == 27 ? 27 : 7;
    constexpr int AL = box_al<TV>();
    constexpr int BCAP = NB + 2 * AL, PCAP = NB + 16;
    constexpr int NCP = box_ncopies(RZ), REC = box_rec_bytes(RZ);
    static_assert(NCP < 32, "one copy per lane of the issuing warp, plus the descriptor");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
    TV* ctab = reinterpret_cast<TV*>(smem_raw + 64);
    TV* dtab = ctab + (size_t)NK * P.NP;
    unsigned char* stage0 = smem_raw + box_head_bytes<TV>(P, NK);
    const unsigned stage_bytes = (unsigned)box_stage_bytes<TV>(P, RZ, NB, NEED_B, NEED_D, NEED_PW);
    const int t = threadIdx.x;
    if (t == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(full + s, 1);
        fence_mbar_init();
    }
    for (int i = t; i < NK * P.NP; i += NB) ctab[i] = ctab_g[i];
    for (int i = t; i < P.NP; i += NB) dtab[i] = ((MODE == 3 && DPAT) || MODE == 4) ? dtab_g[i] : VT<TV>::zero();
    // elements of a window that no copy fills (beyond the ends of the vector) are multiplied by zero coefficients:
    // they must be finite, so the stages start out as zeros
    {
        uint4* z = reinterpret_cast<uint4*>(stage0);
        const unsigned nz = (unsigned)STAGES * stage_bytes / 16;
        for (unsigned i = t; i < nz; i += NB) z[i] = make_uint4(0, 0, 0, 0);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const int ntiles = P.ntiles;
    // Warp 0 issues the copies of a tile: lane i loads entry i of the tile's record and fires it, lane NCP copies the
    // record's descriptor into the stage; lane 0 arms the mbarrier with the byte total first.
    // exchange number this kernel's ghost tiles wait for (read before anything else can advance `consumed`)
    const unsigned long long want = (bw.epoch && t == 0) ? *bw.consumed + 1 : 0;
    bool ghosts_ready = bw.epoch == nullptr;
    auto issue = [&](int tile, int s) {            // all lanes of warp 0
        if (!ghosts_ready && tile >= P.first_ghost_tile) {
            if (t == 0) {
                while (*reinterpret_cast<const volatile unsigned long long*>(bw.epoch) < want) {}
                __threadfence();
            }
            __syncwarp();
            asm volatile("fence.proxy.async.global;" ::: "memory");
            ghosts_ready = true;
        }
        unsigned char* st = stage0 + (size_t)s * stage_bytes;
        const unsigned char* rec = recs + (size_t)tile * REC;
        BoxCopy C;
        C.bytes = 0;
        C.what = -1;
        if (t < NCP) {
            const int4 q = __ldg(reinterpret_cast<const int4*>(rec + BOX_DESC_BYTES) + t);
            C.src = q.x;
            C.dst = (unsigned)q.y;
            C.bytes = (unsigned)q.z;
            C.what = q.w;
            if ((C.what == 1 && !NEED_B) || (C.what == 2 && !NEED_D) || (C.what == 5 && !NEED_PW)) C.bytes = 0;
        } else if (t == NCP) {
            C.what = 4;
            C.dst = 0;
            C.bytes = BOX_DESC_BYTES;
        }
        unsigned total = C.bytes;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
        if (t == 0) mbar_expect_tx(full + s, total);
        __syncwarp();
        if (C.bytes) {
            const void* src = C.what == 0 ? static_cast<const void*>(x + C.src)
                              : (C.what == 1 ? static_cast<const void*>(b + C.src)
                                             : (C.what == 2 ? static_cast<const void*>(d + C.src)
                                                            : ((C.what == 3 || C.what == 5) ? static_cast<const void*>(pid + C.src)
                                                                                            : static_cast<const void*>(rec))));
            bulk_g2s(st + C.dst, src, C.bytes, full + s);
        }
    };
    const bool issuer = t < 32;
    if (issuer && (int)blockIdx.x < ntiles) issue(blockIdx.x, 0);
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int s = STAGES == 2 ? (it & 1) : 0;
        if (STAGES == 2 && issuer && tile + (int)gridDim.x < ntiles) issue(tile + gridDim.x, s ^ 1);
        const unsigned char* st = stage0 + (size_t)s * stage_bytes;
        const BoxDesc* D = reinterpret_cast<const BoxDesc*>(st);
        const TV* sx = reinterpret_cast<const TV*>(st + BOX_DESC_BYTES);
        const uint16_t* sp = reinterpret_cast<const uint16_t*>(sx + P.xtotal);
        const TV* sb = reinterpret_cast<const TV*>(sp + RZ * PCAP);
        const TV* sd = sb + RZ * BCAP;
        const uint16_t* spw = sp + RZ * PCAP;          // MODE 4: the id windows take the place of the b / d tiles
        mbar_wait(full + s, STAGES == 2 ? ((it >> 1) & 1) : (it & 1));
        const int nb = D->nb, nrp = D->nrp;
        const TV* xc[RZ + 2];
        const uint16_t* pwc[RZ + 2];
        const TV* bp[RZ];
        const TV* dp[RZ];
        int pat[RZ];
        bool mine_fast = true;
#pragma unroll
        for (int w = 0; w < RZ + 2; ++w) {
            xc[w] = sx + D->xoff[w] + t;
            pwc[w] = spw + (NEED_PW ? D->pwoff[w] : 0) + t;
        }
#pragma unroll
        for (int j = 0; j < RZ; ++j) {
            bp[j] = sb + D->boff[j] + t;
            dp[j] = sd + D->boff[j] + t;
            const bool ex = t < nb && j < nrp;
            pat[j] = ex ? (int)sp[D->poff[j] + t] : P.p0;
            mine_fast = mine_fast && (pat[j] == P.p0);
        }
        const bool fast = __all_sync(0xffffffffu, mine_fast);
        if (t < nb) {
            TV out[RZ];
            if (fast) box_thread<TV, SHAPE, MODE, DPAT, RZ, true>(C0, ctab, dtab, P.NP, P.S, xc, bp, dp, pat, out, pwc);
            else box_thread<TV, SHAPE, MODE, DPAT, RZ, false>(C0, ctab, dtab, P.NP, P.S, xc, bp, dp, pat, out, pwc);
            const int row0 = D->r0 + t;
#pragma unroll
            for (int j = 0; j < RZ; ++j)
                if (j < nrp) {
                    const int row = row0 + j * P.S2;
                    y[row] = out[j];
                    if (pp.on) ll_put_edge<TV>(pp, row, out[j]);
                }
        }
        __syncthreads();
        if (STAGES == 1 && issuer && tile + (int)gridDim.x < ntiles) issue(tile + gridDim.x, 0);
    }
    if (bw.epoch) {          // the last CTA to finish marks the exchange as consumed
        __syncthreads();
        if (t == 0 && atomicInc(bw.ticket, gridDim.x - 1) == gridDim.x - 1) *bw.consumed = want;
    }
}

// ---- direct form: no staging at all ----------------------------------------------------------------------------------
// The same unrolled chains with x read straight from global memory through the read-only path: every load of a warp is
// a contiguous 256-byte segment, the +-1 and +-S neighbours hit L1, the +-S2 planes hit L2, and nothing synchronises -
// 64 independent warps per SM with 7 ... 27 loads in flight each hide the latency that the staged form has to hide
// with its two-stage pipeline and a CTA barrier per tile (profiles/r02e_ncu_box_v0_raw.csv: barrier stall 13.7 cycles
// per issue, no pipe above 47 %).  Rows whose zero-coefficient neighbours would lie outside the input vector (the first
// and the last plane) walk the dictionary entry by entry instead, like pat_kernel.
template <typename TV, int SHAPE, int MODE, bool DPAT, int RZ, int NT>
__global__ void __launch_bounds__(NT)
box_direct_kernel(const __grid_constant__ BoxPlan P, const __grid_constant__ BoxCoef<TV> C0, const __grid_constant__ PutPlan pp,
                  const uint16_t* __restrict__ pid, const TV* __restrict__ ctab_g, const TV* __restrict__ dtab_g,
                  const int* __restrict__ pat_off, const PatEntry<TV>* __restrict__ ent, const TV* __restrict__ x,
                  const TV* __restrict__ b, const TV* __restrict__ d, TV* __restrict__ y) {
    constexpr int NK = SHAPE == 27 ? 27 : 7;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TV* ctab = reinterpret_cast<TV*>(smem_raw);
    TV* dtab = ctab + (size_t)NK * P.NP;
    const int t = threadIdx.x;
    for (int i = t; i < NK * P.NP; i += NT) ctab[i] = ctab_g[i];
    for (int i = t; i < P.NP; i += NT) dtab[i] = (MODE == 3 && DPAT) ? dtab_g[i] : VT<TV>::zero();
    __syncthreads();
    // plane groups of RZ planes in blockIdx.y (RZ == 1: one "plane" of n_rows), chunks of the plane in blockIdx.x
    const int g = blockIdx.y;
    const int nrp = P.nplanes - g * RZ < RZ ? P.nplanes - g * RZ : RZ;
    const int gbase = g * RZ * P.plane;
    for (int c0 = blockIdx.x * NT; c0 < P.plane; c0 += gridDim.x * NT) {       // warp-uniform trip count
        const int c = c0 + t;
        const bool act = c < P.plane;
        const int r0 = gbase + (act ? c : 0);
        int pat[RZ];
        bool mine_fast = true;
#pragma unroll
        for (int j = 0; j < RZ; ++j) {
            pat[j] = (act && j < nrp) ? (int)__ldg(reinterpret_cast<const unsigned short*>(pid) + r0 + j * P.S2) : P.p0;
            mine_fast = mine_fast && (pat[j] == P.p0);
        }
        // every address of the unrolled chain inside the input vector?  (windows 0 .. RZ+1, offsets up to +-(S+1))
        const bool safe = nrp == RZ && r0 - P.S2 - P.S - 1 >= P.xlo && r0 + RZ * P.S2 + P.S + 1 < P.xhi;
        const bool fast = __all_sync(0xffffffffu, mine_fast && (safe || !act));
        if (!act) continue;
        if (safe) {
            const TV* xc[RZ + 2];
            const TV* bp[RZ];
            const TV* dp[RZ];
#pragma unroll
            for (int w = 0; w < RZ + 2; ++w) xc[w] = x + r0 + (w - 1) * P.S2;
#pragma unroll
            for (int j = 0; j < RZ; ++j) {
                bp[j] = b + r0 + j * P.S2;
                dp[j] = d + r0 + j * P.S2;
            }
            TV out[RZ];
            if (fast) box_thread<TV, SHAPE, MODE, DPAT, RZ, true>(C0, ctab, dtab, P.NP, P.S, xc, bp, dp, pat, out);
            else box_thread<TV, SHAPE, MODE, DPAT, RZ, false>(C0, ctab, dtab, P.NP, P.S, xc, bp, dp, pat, out);
#pragma unroll
            for (int j = 0; j < RZ; ++j) {
                const int row = r0 + j * P.S2;
                y[row] = out[j];
                if (pp.on) ll_put_edge<TV>(pp, row, out[j]);
            }
        } else {
            for (int j = 0; j < nrp; ++j) {          // exact walk: only the entries the row has
                const int row = r0 + j * P.S2;
                const int k0 = __ldg(pat_off + pat[j]), k1 = __ldg(pat_off + pat[j] + 1);
                TV acc = VT<TV>::zero();
                for (int k = k0; k < k1; ++k) {
                    const PatEntry<TV> e = ldg_ent(ent + k);
                    acc = acc + e.v * ldg_(x + (row + e.delta));
                }
                TV bval = VT<TV>::zero(), dval = VT<TV>::zero(), xval = VT<TV>::zero();
                if (MODE == 2 || MODE == 3) bval = b[row];
                if (MODE == 3) {
                    dval = DPAT ? dtab[pat[j]] : d[row];
                    xval = x[row];
                }
                const TV out = pat_epilogue<MODE, TV>(acc, xval, bval, dval);
                y[row] = out;
                if (pp.on) ll_put_edge<TV>(pp, row, out);
            }
        }
    }
}

// ---- block variant (nrhs > 1): the same unrolled chains over a block of right-hand sides ------------------------------------
// Blocks are stored RHS-fastest (x[row * m + j], solver.cuh), so the m values of a row are one contiguous segment: a
// group of MP = min(32, pow2(m)) lanes owns one row, lane j its right-hand side j (+ MP, + 2 MP ... when m > 32), and every
// load of a group is one coalesced segment (256 bytes for m = 32).  The matrix is not streamed at all - pattern id per
// row, coefficients from the constant bank (groups of a warp that all carry the dominant pattern) or from the dense
// table - where csr_stream_mrhs_kernel reads 12 bytes per non-zero.  Products run in stored order per (row, j):
// bit-identical to the CSR kernel and to the oracle's column-by-column SpMatMul.
template <typename TV, int SHAPE, int MODE, bool DPAT, int NT>
__global__ void __launch_bounds__(NT)
box_mrhs_kernel(const __grid_constant__ BoxPlan P, const __grid_constant__ BoxCoef<TV> C0, int m, int mp,
                const uint16_t* __restrict__ pid, const TV* __restrict__ ctab_g, const TV* __restrict__ dtab_g,
                const int* __restrict__ pat_off, const PatEntry<TV>* __restrict__ ent, const TV* __restrict__ x,
                const TV* __restrict__ b, const TV* __restrict__ d, TV* __restrict__ y) {
    constexpr int NK = SHAPE == 27 ? 27 : 7;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TV* ctab = reinterpret_cast<TV*>(smem_raw);
    TV* dtab = ctab + (size_t)NK * P.NP;
    const int t = threadIdx.x;
    for (int i = t; i < NK * P.NP; i += NT) ctab[i] = ctab_g[i];
    for (int i = t; i < P.NP; i += NT) dtab[i] = (MODE == 3 && DPAT) ? dtab_g[i] : VT<TV>::zero();
    __syncthreads();
    const int rpb = NT / mp;                       // rows per CTA pass
    const int rloc = t / mp, j0 = t - rloc * mp;
    const long long S = P.S, S2 = P.S2;
    for (int base = blockIdx.x * rpb; base < P.n_rows; base += gridDim.x * rpb) {       // warp-uniform trip count
        const int row = base + rloc;
        const bool act = row < P.n_rows;
        const int pat = act ? (int)__ldg(reinterpret_cast<const unsigned short*>(pid) + row) : P.p0;
        const bool safe = row - S2 - S - 1 >= P.xlo && row + S2 + S + 1 < P.xhi;
        const bool fast = __all_sync(0xffffffffu, !act || (pat == P.p0 && safe));
        if (!act) continue;
        const TV dval = (MODE == 3) ? (DPAT ? dtab[pat] : d[row]) : VT<TV>::zero();
        for (int j = j0; j < m; j += mp) {
            const TV* xr = x + (long long)row * m + j;
            TV acc = VT<TV>::zero();
            if (safe) {
#pragma unroll
                for (int dz = -1; dz <= 1; ++dz)
#pragma unroll
                    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
                        for (int dx = -1; dx <= 1; ++dx) {
                            if (!box_has<SHAPE>(dz, dy, dx)) continue;
                            const int k = box_k<SHAPE>(dz, dy, dx);
                            const TV cf = fast ? C0.c[k] : ctab[k * P.NP + pat];
                            acc = acc + cf * ldg_(xr + (dz * S2 + dy * S + dx) * m);
                        }
            } else {                                // rows next to the ends of the vector: only the entries the row has
                const int k0 = __ldg(pat_off + pat), k1 = __ldg(pat_off + pat + 1);
                for (int k = k0; k < k1; ++k) {
                    const PatEntry<TV> e = ldg_ent(ent + k);
                    acc = acc + e.v * ldg_(xr + (long long)e.delta * m);
                }
            }
            TV bval = VT<TV>::zero(), xval = VT<TV>::zero();
            if (MODE == 2 || MODE == 3) bval = b[(long long)row * m + j];
            if (MODE == 3) xval = *xr;
            y[(long long)row * m + j] = pat_epilogue<MODE, TV>(acc, xval, bval, dval);
        }
    }
}

// ---- block variant, marching form ------------------------------------------------------------------------------------------
// box_mrhs_kernel above reads every neighbour of a row from global memory, and the CTAs that run together are spread over
// a whole plane of the grid: nothing is reused in L1, each block row of x is read 7 (27) times from L2, and the kernel
// runs at the L2's pace (cfg4, profiles/r02o_cfg4.json: level-1 sweep 612 us for 1.65 GB of vectors, 8 TB/s through
// L2).  Here a CTA owns a tile of TX x TY grid columns and marches through a range of planes (2.5-D blocking): a node's
// m values (256 bytes for m = 32) enter shared memory once per CTA - four rotating plane buffers of (TX + 2) x (TY + 2)
// nodes - and all 7 / 27 neighbours are read from there; one warp = one node at a time, lane j = right-hand side j,
// every shared-memory access a conflict-free 256-byte row, one barrier per plane, and the global loads of plane k + 2
// fly while plane k is computed.  Nodes outside the grid are zeros in the buffers (their coefficients are zero in the
// dense table, box_kernel's argument).  Products run in stored (dz, dy, dx) order per (row, j): bit-identical to
// box_mrhs_kernel, the CSR block kernel and the oracle's column-by-column SpMatMul.
// (A form that also carried b and the pattern ids one plane ahead and issued all loads before the barrier measured
// twice as slow - 1219 against 598 us on the level-1 sweep of cfg4, profiles/r02u_tune_cfg4.log - and was dropped.)
constexpr int MARCH_TX = 8, MARCH_TY = 8, MARCH_NW = 16;       // tile columns, warps per CTA
constexpr int MARCH_HX = MARCH_TX + 2, MARCH_HY = MARCH_TY + 2, MARCH_NODES = MARCH_HX * MARCH_HY;
constexpr int MARCH_HALO = MARCH_NODES - MARCH_TX * MARCH_TY;   // 36 nodes around the tile
template <typename TV>
static inline size_t march_smem_bytes(int shape, int NP) {
    return ((size_t)shape * NP + NP) * sizeof(TV) + 4 * (size_t)MARCH_NODES * 32 * sizeof(TV);
}
template <typename TV, int SHAPE, int MODE, bool DPAT>
__global__ void __launch_bounds__(MARCH_NW * 32, 2)
box_mrhs_march_kernel(const __grid_constant__ BoxCoef<TV> C0, int NP, int p0, int m, int n0, int n1, int nz, int zchunk,
                      const uint16_t* __restrict__ pid, const TV* __restrict__ ctab_g, const TV* __restrict__ dtab_g,
                      const TV* __restrict__ x, const TV* __restrict__ b, const TV* __restrict__ d, TV* __restrict__ y) {
    constexpr int NK = SHAPE == 27 ? 27 : 7;
    constexpr int TX = MARCH_TX, TY = MARCH_TY, HX = MARCH_HX, NODES = MARCH_NODES, NW = MARCH_NW;
    constexpr int CPW = TX * TY / NW;                 // columns per warp (consecutive in x)
    constexpr int HPW = (MARCH_HALO + NW - 1) / NW;   // halo nodes per warp
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TV* ctab = reinterpret_cast<TV*>(smem_raw);
    TV* dtab = ctab + (size_t)NK * NP;
    TV* buf = dtab + NP;                              // [4][NODES][32]
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    for (int i = t; i < NK * NP; i += NW * 32) ctab[i] = ctab_g[i];
    for (int i = t; i < NP; i += NW * 32) dtab[i] = (MODE == 3 && DPAT) ? dtab_g[i] : VT<TV>::zero();
    const int tiles_x = (n0 + TX - 1) / TX;
    const int tx0 = ((int)blockIdx.x % tiles_x) * TX, ty0 = ((int)blockIdx.x / tiles_x) * TY;
    const int kz0 = (int)blockIdx.y * zchunk, kz1 = min(nz, kz0 + zchunk);
    const int j = (int)blockIdx.z * 32 + lane;
    const bool jl = j < m;
    const long long S2 = (long long)n0 * n1;
    // the nodes this warp moves into the plane buffers: CPW tile columns and up to HPW halo nodes
    int node[CPW + HPW];          // index in a plane buffer, -1: none
    int col[CPW + HPW];           // gy * S + gx, -1: outside the grid
#pragma unroll
    for (int q = 0; q < CPW; ++q) {
        const int c = w * CPW + q, lx = c % TX, ly = c / TX, gx = tx0 + lx, gy = ty0 + ly;
        node[q] = (ly + 1) * HX + (lx + 1);
        col[q] = (gx < n0 && gy < n1) ? gy * n0 + gx : -1;
    }
#pragma unroll
    for (int q = 0; q < HPW; ++q) {
        const int h = w + q * NW;
        int hx, hy;
        if (h < HX) { hy = -1; hx = h - 1; }
        else if (h < 2 * HX) { hy = TY; hx = h - HX - 1; }
        else if (h < 2 * HX + TY) { hx = -1; hy = h - 2 * HX; }
        else { hx = TX; hy = h - 2 * HX - TY; }
        const int gx = tx0 + hx, gy = ty0 + hy;
        node[CPW + q] = h < MARCH_HALO ? (hy + 1) * HX + (hx + 1) : -1;
        col[CPW + q] = (h < MARCH_HALO && gx >= 0 && gx < n0 && gy >= 0 && gy < n1) ? gy * n0 + gx : -1;
    }
    TV nx[CPW + HPW];
    auto load_plane = [&](int kk) {
#pragma unroll
        for (int q = 0; q < CPW + HPW; ++q) {
            nx[q] = VT<TV>::zero();
            if (kk >= 0 && kk < nz && col[q] >= 0 && jl) nx[q] = ldg_(x + ((kk * S2 + col[q]) * m + j));
        }
    };
    auto store_plane = [&](int kk) {
        TV* pb = buf + (size_t)(kk & 3) * NODES * 32;
#pragma unroll
        for (int q = 0; q < CPW + HPW; ++q)
            if (node[q] >= 0) pb[node[q] * 32 + lane] = nx[q];
    };
    load_plane(kz0 - 1);
    store_plane(kz0 - 1);
    load_plane(kz0);
    store_plane(kz0);
    load_plane(kz0 + 1);
    for (int k = kz0; k < kz1; ++k) {
        store_plane(k + 1);
        __syncthreads();
        load_plane(k + 2);                       // in flight while plane k is computed
        const TV* pz[3] = {buf + (size_t)((k - 1) & 3) * NODES * 32 + lane, buf + (size_t)(k & 3) * NODES * 32 + lane,
                           buf + (size_t)((k + 1) & 3) * NODES * 32 + lane};
        TV bv[CPW];
        int pat[CPW];
#pragma unroll
        for (int q = 0; q < CPW; ++q) {
            bv[q] = VT<TV>::zero();
            pat[q] = p0;
            if (col[q] >= 0) {
                const long long row = k * S2 + col[q];
                pat[q] = (int)__ldg(reinterpret_cast<const unsigned short*>(pid) + row);
                if ((MODE == 2 || MODE == 3) && jl) bv[q] = b[row * m + j];
            }
        }
#pragma unroll
        for (int q = 0; q < CPW; ++q) {
            if (col[q] < 0) continue;              // warp-uniform
            const long long row = k * S2 + col[q];
            const bool fast = pat[q] == p0;        // warp-uniform
            TV acc = VT<TV>::zero();
#pragma unroll
            for (int dz = -1; dz <= 1; ++dz)
#pragma unroll
                for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
                    for (int dx = -1; dx <= 1; ++dx) {
                        if (!box_has<SHAPE>(dz, dy, dx)) continue;
                        const int kk = box_k<SHAPE>(dz, dy, dx);
                        const TV cf = fast ? C0.c[kk] : ctab[kk * NP + pat[q]];
                        acc = acc + cf * pz[dz + 1][(node[q] + dy * HX + dx) * 32];
                    }
            TV dval = VT<TV>::zero(), xval = VT<TV>::zero();
            if (MODE == 3) {
                dval = DPAT ? (fast ? C0.d0 : dtab[pat[q]]) : d[row];
                xval = pz[1][node[q] * 32];
            }
            if (jl) y[row * m + j] = pat_epilogue<MODE, TV>(acc, xval, bv[q], dval);
        }
    }
}

// ---- host side: dense tables ---------------------------------------------------------------------------------------
template <typename TV>
struct BoxDict {
    bool ok = false;
    int shape = 0;                 // 7 or 27
    int npat = 0, NP = 0, p0 = 0;
    BoxCoef<TV> c0;
    TV* ctab = nullptr;            // device: coef[k * NP + p]
    TV* dtab = nullptr;            // device: folded relaxation weights per pattern (set by fold_d), NP elements
    std::vector<TV> h_ctab;        // host copies (the CPU replay, tests)
    std::vector<int> h_mask, h_pat_off;   // presence bits and entry offsets of the patterns (refill)
    // tile records (box_plan_tile) per kernel variant in use - the fused first two sweeps may run another variant than the
    // other modes - planned at the first launch of a variant and whenever the copyable range of the input vectors changes
    struct RecSet {
        unsigned char* recs = nullptr;
        int RZ = 0, NB = 0, xlo = 0, xhi = 0, first_ghost_tile = 0;
    };
    std::vector<RecSet> sets;
    void release() {
        if (ctab) cudaFree(ctab);
        if (dtab) cudaFree(dtab);
        for (auto& r : sets)
            if (r.recs) cudaFree(r.recs);
        sets.clear();
        ctab = dtab = nullptr;
        ok = false;
        h_ctab.clear();
        h_mask.clear();
        h_pat_off.clear();
    }
    // the same tables from new dictionary values (replace_matrix: the patterns kept their shape, galerkin.cuh)
    bool refill(const std::vector<TV>& val) {
        if (!ok || h_mask.empty() || (int)h_pat_off.size() != npat + 1) return false;
        std::fill(h_ctab.begin(), h_ctab.end(), VT<TV>::zero());
        for (int p = 0; p < npat; ++p) {
            int e = h_pat_off[p];
            for (int bit = 0; bit < 27; ++bit) {
                if (!(h_mask[p] & (1 << bit))) continue;
                const int dz = bit / 9 - 1, dy = (bit / 3) % 3 - 1, dx = bit % 3 - 1;
                const int k = shape == 27 ? box_k<27>(dz, dy, dx) : box_k<7>(dz, dy, dx);
                h_ctab[(size_t)k * NP + p] = val[e++];
            }
            if (e != h_pat_off[p + 1]) return false;
        }
        for (int k = 0; k < shape; ++k) c0.c[k] = h_ctab[(size_t)k * NP + p0];
        return true;
    }
    const RecSet* find_records(const BoxPlan& P, int RZ, int NB) const {
        for (const auto& r : sets)
            if (r.recs && r.RZ == RZ && r.NB == NB && r.xlo == P.xlo && r.xhi == P.xhi) return &r;
        return nullptr;
    }
    bool has_records(const BoxPlan& P, int RZ, int NB) const { return find_records(P, RZ, NB) != nullptr; }
    int first_ghost_tile = 0;          // of the set records() returned last
    const unsigned char* records(const BoxPlan& P, int RZ, int NB) {
        if (const RecSet* r = find_records(P, RZ, NB)) {
            first_ghost_tile = r->first_ghost_tile;
            return r->recs;
        }
        // a set planned for another range of the same variant is stale: drop it
        for (auto& r : sets)
            if (r.recs && r.RZ == RZ && r.NB == NB) {
                cudaFree(r.recs);
                r.recs = nullptr;
            }
        const size_t rb = box_rec_bytes(RZ);
        std::vector<unsigned char> h((size_t)P.ntiles * rb), one(rb);
        // tiles that read ghost rows of the input vector (rows outside [0, n_rows)) go last: on a row-partitioned level
        // they may have to wait for the halo exchange that runs beside the kernel
        constexpr int AL = box_al<TV>();
        const int nal = box_ceil(P.n_rows, AL);
        std::vector<int> ghost_tiles;
        int n_int = 0;
        for (int tile = 0; tile < P.ntiles; ++tile) {
            box_plan_tile<TV>(P, RZ, NB, tile, one.data());
            const BoxCopy* cp = reinterpret_cast<const BoxCopy*>(one.data() + BOX_DESC_BYTES);
            bool ghost = false;
            for (int i = 0; i < RZ + 2; ++i)
                if (cp[i].bytes && (cp[i].src < 0 || cp[i].src + (int)(cp[i].bytes / sizeof(TV)) > nal)) ghost = true;
            if (ghost) ghost_tiles.push_back(tile);
            else std::memcpy(h.data() + (size_t)(n_int++) * rb, one.data(), rb);
        }
        RecSet R;
        R.first_ghost_tile = n_int;
        for (int tile : ghost_tiles) box_plan_tile<TV>(P, RZ, NB, tile, h.data() + (size_t)(n_int++) * rb);
        MGB_CUDA(cudaMalloc(&R.recs, std::max<size_t>(h.size(), 16)));
        MGB_CUDA(cudaMemcpy(R.recs, h.data(), h.size(), cudaMemcpyHostToDevice));
        R.RZ = RZ;
        R.NB = NB;
        R.xlo = P.xlo;
        R.xhi = P.xhi;
        bool placed = false;
        for (auto& r : sets)
            if (!r.recs) {
                r = R;
                placed = true;
                break;
            }
        if (!placed) sets.push_back(R);
        first_ghost_tile = R.first_ghost_tile;
        return R.recs;
    }
};
// dense table of a box-structured dictionary; false when the shape is not one the kernel is instantiated for
template <typename TV>
static bool box_build_tables(const HostPatterns<TV>& H, const BoxInfo& B, long long n_rows, int& shape, int& NP, int& p0,
                             std::vector<TV>& tab, BoxCoef<TV>& c0) {
    if (!B.ok || B.S < 3 || B.S2 < 3 * B.S || H.npat() > BOX_MAX_PAT) return false;
    int U = 0;
    for (int m : B.mask) U |= m;
    shape = (U & ~BOX_STAR_MASK) == 0 ? 7 : 27;
    const int nk = shape;
    NP = (H.npat() + 1) & ~1;
    tab.assign((size_t)nk * NP, VT<TV>::zero());
    for (int p = 0; p < H.npat(); ++p) {
        int e = H.pat_off[p];
        for (int bit = 0; bit < 27; ++bit) {
            if (!(B.mask[p] & (1 << bit))) continue;
            const int dz = bit / 9 - 1, dy = (bit / 3) % 3 - 1, dx = bit % 3 - 1;
            const int k = shape == 27 ? box_k<27>(dz, dy, dx) : box_k<7>(dz, dy, dx);
            tab[(size_t)k * NP + p] = H.val[e++];
        }
        if (e != H.pat_off[p + 1]) return false;
    }
    std::vector<long long> cnt(H.npat(), 0);
    for (long long r = 0; r < n_rows; ++r) cnt[H.pid[r]]++;
    p0 = (int)(std::max_element(cnt.begin(), cnt.end()) - cnt.begin());
    std::memset(static_cast<void*>(&c0), 0, sizeof(c0));
    for (int k = 0; k < nk; ++k) c0.c[k] = tab[(size_t)k * NP + p0];
    return true;
}

}  // namespace mgb200
