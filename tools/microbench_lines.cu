// Micro-benchmark for the NEXT form of the stencil-dictionary sweep: reuse across rows.
//
// The production kernel (csrc/pattern.cuh, pat_tma_kernel) gives one row to one thread and is bound by shared-memory
// wavefronts: per stencil entry a value (2 wavefronts), an offset (1) and x (2).  Here a thread owns R rows of the SAME
// column of R consecutive x-lines (rows i + j*S, S = line length; lanes are contiguous in i, so every load of a warp is
// contiguous).  Per z-plane of the stencil it loads x[dx][l], dx in {-1,0,1}, l in [-1, R], once - 3(R+2) loads for
// 9R products - and every dictionary value once for R rows.  The offsets are not read at all: a pattern is a 27-bit
// presence mask over (dz,dy,dx) plus its values in stored order, and stored order IS (dz,dy,dx) order, so every row
// still accumulates its products in stored order: results are bit-identical to the one-row-per-thread kernel.
// Groups whose R rows do not share one pattern (first / last lines of a plane when R does not divide the line count)
// fall back to one row at a time.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -lineinfo -o gpurun_out/microbench_lines tools/microbench_lines.cu
//   gpurun_out/microbench_lines              # 7-point 257^3 and 27-point 129^3 on the GPU, all variants, bit-compare
//   gpurun_out/microbench_lines --host-check # no GPU: runs the per-thread function on the CPU for 9^3 .. 12x9x7 grids
//
// Three forms: (a) x straight from global memory (L1 / L2), (b) x staged in shared memory by bulk copies, one per plane of
// the stencil, in a two-stage pipeline of persistent CTAs - the production kernel's scheme with another thread-to-row map,
// (c) marching: a CTA walks a column of line groups plane by plane with a ring of four plane slabs, so every x element
// enters shared memory once per column instead of once per stencil plane (2.5-D blocking).
//
// NOT part of the library; not yet run on a B200 (written in a session whose GPU budget was spent; the host check,
// which also replays the staged form with host buffers filled like the bulk copies fill shared memory, passes).
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

struct __align__(16) Ent { double v; int delta; int pad; };
struct Pat { int k0, len, mask; };   // entries [k0, k0+len), presence mask bit (dz+1)*9 + (dy+1)*3 + (dx+1)

struct Grid { int n1, n2, n3, S, S2; long long n; };   // S = n1 (line), S2 = n1*n2 (plane)

// reference: one row per thread, dictionary walked entry by entry (what pat_kernel does)
__host__ __device__ inline double row_reference(const Grid& G, long long row, const uint16_t* pid, const Pat* pat,
                                                const Ent* ent, const double* dp, const double* x, const double* b) {
    const int p = pid[row];
    const Pat P = pat[p];
    double acc = 0.0;
    for (int k = P.k0; k < P.k0 + P.len; ++k) acc = acc + ent[k].v * x[row + ent[k].delta];
    return x[row] + dp[p] * (b[row] - acc);
}
__global__ void __launch_bounds__(256) k_reference(Grid G, const uint16_t* __restrict__ pid, const Pat* __restrict__ pat,
                                                   const Ent* __restrict__ ent, const double* __restrict__ dp,
                                                   const double* __restrict__ x, const double* __restrict__ b,
                                                   double* __restrict__ y) {
    const long long row = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (row < G.n) y[row] = row_reference(G, row, pid, pat, ent, dp, x, b);
}

// Where a thread finds its data.  rel = row - T0 for a reference row T0 (0 for the global-memory form, the first row of
// the tile for the staged form).  xs[dz+1][rel] == x[T0 + dz*S2 + rel] for every rel the thread may touch.
struct View {
    const double* xs[3];
    const double* b;        // b[rel]
    const uint16_t* pid;    // pid[rel]
};

// fallback: one row, dictionary walked entry by entry, x through the view (delta = dx + S*dy + S2*dz)
__host__ __device__ inline double row_view(const Grid& G, const View& V, long long rel, const Pat* pat, const Ent* ent,
                                           const double* dp) {
    const int p = V.pid[rel];
    const Pat P = pat[p];
    double acc = 0.0;
    for (int k = P.k0; k < P.k0 + P.len; ++k) {
        const int d = ent[k].delta;
        const int dz = (2 * d > G.S2) ? 1 : ((2 * d < -G.S2) ? -1 : 0);
        const double* xb = dz > 0 ? V.xs[2] : (dz < 0 ? V.xs[0] : V.xs[1]);      // no dynamic index into the view
        acc = acc + ent[k].v * xb[rel + (d - dz * G.S2)];
    }
    return V.xs[1][rel] + dp[p] * (V.b[rel] - acc);
}

// one thread: column i of the R lines [y0, y0+R) of plane z; rel0 = rel of its first row, row0 = that row's global index
template <int R>
__host__ __device__ inline void lines_thread(const Grid& G, const View& V, long long rel0, long long row0, int y0,
                                             const Pat* pat, const Ent* ent, const double* dp, double* y) {
    const int nr = (G.n2 - y0 < R) ? (G.n2 - y0) : R;          // lines left in this plane
    const int p0 = V.pid[rel0];
    bool same = (nr == R);
#pragma unroll
    for (int j = 1; j < R; ++j)
        if (j < nr) same = same && (V.pid[rel0 + (long long)j * G.S] == p0);
    if (!same) {                                                  // mixed patterns or a short group: row by row
        for (int j = 0; j < nr; ++j) y[row0 + (long long)j * G.S] = row_view(G, V, rel0 + (long long)j * G.S, pat, ent, dp);
        return;
    }
    const Pat P = pat[p0];
    const Ent* e = ent + P.k0;
    double acc[R], xc[R];
    bool have_c = false;                                          // x[row] comes along with the dz = 0 plane, normally
#pragma unroll
    for (int j = 0; j < R; ++j) { acc[j] = 0.0; xc[j] = 0.0; }
#pragma unroll
    for (int dz = -1; dz <= 1; ++dz) {
        const int mz = (P.mask >> ((dz + 1) * 9)) & 0x1FF;
        if (mz == 0) continue;
        const double* xp = V.xs[dz + 1] + rel0;
        // X[dx][l] is loaded iff some row of the group multiplies it: row j uses line l = j + 1 + dy, so dy = -1 reaches
        // l <= R-1, dy = 0 the lines 1..R and dy = +1 l >= 2.  Every load is then an address some row's entry names,
        // hence inside the vector - the kernel assumes nothing about the grid beyond the offsets dz*S2 + dy*S + dx.
        const bool col_0 = (mz & 0x092) != 0;
        double X[3][R + 2];
#pragma unroll
        for (int l = 0; l < R + 2; ++l) {
            const double* q = xp + (long long)(l - 1) * G.S;
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
                const bool need = ((l <= R - 1) && (mz & (1 << (0 + dx + 1)))) || ((l >= 1 && l <= R) && (mz & (1 << (3 + dx + 1)))) ||
                                  ((l >= 2) && (mz & (1 << (6 + dx + 1))));
                X[dx + 1][l] = need ? q[dx] : 0.0;
            }
        }
        if (dz == 0 && col_0) {
            have_c = true;
#pragma unroll
            for (int j = 0; j < R; ++j) xc[j] = X[1][j + 1];
        }
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
                if (mz & (1 << ((dy + 1) * 3 + (dx + 1)))) {
                    const double v = e->v;
                    ++e;
#pragma unroll
                    for (int j = 0; j < R; ++j) acc[j] = acc[j] + v * X[dx + 1][j + 1 + dy];
                }
            }
        }
    }
    const double d = dp[p0];
#pragma unroll
    for (int j = 0; j < R; ++j) {
        const long long rel = rel0 + (long long)j * G.S;
        if (!have_c) xc[j] = V.xs[1][rel];
        y[row0 + (long long)j * G.S] = xc[j] + d * (V.b[rel] - acc[j]);
    }
}

__host__ __device__ inline View global_view(const Grid& G, const uint16_t* pid, const double* x, const double* b) {
    View V;
    V.xs[0] = x - G.S2; V.xs[1] = x; V.xs[2] = x + G.S2;
    V.b = b; V.pid = pid;
    return V;
}

// (a) straight from global memory: persistent grid-stride over the flattened (group, column) index; groups ordered plane
// by plane so that the three planes a group reads stay L2 resident.  x must be addressable S2 elements beyond both ends
// (only the pointer arithmetic goes there, no load does).
template <int R>
__global__ void __launch_bounds__(256) k_lines(Grid G, const uint16_t* __restrict__ pid, const Pat* __restrict__ pat,
                                               const Ent* __restrict__ ent, const double* __restrict__ dp,
                                               const double* __restrict__ x, const double* __restrict__ b,
                                               double* __restrict__ y) {
    const int gpp = (G.n2 + R - 1) / R;                       // line groups per plane
    const long long total = (long long)G.n3 * gpp * G.S;
    const View V = global_view(G, pid, x, b);
    for (long long f = blockIdx.x * (long long)blockDim.x + threadIdx.x; f < total; f += (long long)gridDim.x * blockDim.x) {
        const long long g = f / G.S;
        const int i = (int)(f - g * G.S);
        const int z = (int)(g / gpp), q = (int)(g - (long long)z * gpp);
        const long long row0 = (long long)z * G.S2 + (long long)q * R * G.S + i;
        lines_thread<R>(G, V, row0, row0, q * R, pat, ent, dp, y);
    }
}

// (b) staged: a tile is Q line groups of one plane (Q*R lines = Q*R*S consecutive rows).  Its x comes as three
// contiguous ranges, one per plane of the stencil, each ONE bulk copy (the merged windows of csrc/pattern.cuh with a
// tile of Q*R*S rows), plus the b and pid tiles.  Ranges are clipped to the vector and rounded outwards to 16 bytes.
struct TilePlan {
    long long T0;            // first row of the tile
    int lines;               // lines in the tile (<= Q*R)
    long long xa[3], xe[3];  // copy range of x per stencil plane, [xa, xe), even; xe <= xa: plane absent
    long long ba, be;        // copy range of b (even)
    long long pa, pe;        // copy range of pid (multiples of 8)
};
__host__ __device__ inline TilePlan plan_tile(const Grid& G, int R, int Q, long long tile) {
    const int gpp = (G.n2 + R - 1) / R, tpp = (gpp + Q - 1) / Q;
    const int z = (int)(tile / tpp), q0 = (int)(tile - (long long)z * tpp) * Q;
    TilePlan T;
    T.T0 = (long long)z * G.S2 + (long long)q0 * R * G.S;
    const int left = G.n2 - q0 * R;
    T.lines = left < Q * R ? left : Q * R;
    const long long T1 = T.T0 + (long long)T.lines * G.S, nev = (G.n + 1) & ~1LL, n8 = (G.n + 7) & ~7LL;
#pragma unroll
    for (int dz = -1; dz <= 1; ++dz) {
        if (z + dz < 0 || z + dz >= G.n3) { T.xa[dz + 1] = T.xe[dz + 1] = 0; continue; }
        long long a = T.T0 + (long long)dz * G.S2 - G.S - 1, e = T1 + (long long)dz * G.S2 + G.S + 1;
        a = a < 0 ? 0 : a;
        e = e > nev ? nev : e;
        T.xa[dz + 1] = a & ~1LL;
        T.xe[dz + 1] = (e + 1) & ~1LL;
    }
    T.ba = T.T0 & ~1LL;
    T.be = ((T1 + 1) & ~1LL) > nev ? nev : ((T1 + 1) & ~1LL);
    T.pa = T.T0 & ~7LL;
    T.pe = ((T1 + 7) & ~7LL) > n8 ? n8 : ((T1 + 7) & ~7LL);
    return T;
}
// elements a stage must hold (same for every tile): 3 x-ranges, b, pid
__host__ __device__ inline void stage_layout(const Grid& G, int R, int Q, int& xcap, int& bcap, int& pcap) {
    xcap = (Q * R + 2) * G.S + 6;          // + 2 (the +-1 columns) + rounding
    bcap = Q * R * G.S + 4;
    pcap = Q * R * G.S + 16;
    xcap = (xcap + 1) & ~1; bcap = (bcap + 1) & ~1; pcap = (pcap + 7) & ~7;
}
__host__ __device__ inline View stage_view(const Grid& G, const TilePlan& T, const double* sx, int xcap, const double* sb,
                                           const uint16_t* sp) {
    View V;
    // xs[d][rel] = stage_d[(T0 + (d-1)*S2 + rel) - xa[d]]
#pragma unroll
    for (int d = 0; d < 3; ++d) V.xs[d] = sx + (long long)d * xcap + ((T.T0 + (long long)(d - 1) * G.S2) - T.xa[d]);
    V.b = sb + (T.T0 - T.ba);
    V.pid = sp + (T.T0 - T.pa);
    return V;
}

// (c) marching: a CTA owns a column of Q line groups and walks it plane by plane (2.5-D blocking).  Slot zz % 4 of a
// ring holds, for plane zz, the x slab [first line - 1, last line + 1] of the column (one bulk copy) and - for planes
// the CTA computes - the b and pid tiles.  Computing plane z reads the slabs of z-1, z, z+1 while z+2 is in flight:
// every x element enters shared memory ONCE per column (plus the two halo lines), not once per stencil plane.
struct SlabPlan {
    long long T0;            // first row of the column in this plane
    int lines;
    long long xa, xe;        // x slab [xa, xe), even; xe <= xa: plane outside the grid
    long long ba, be, pa, pe;
};
__host__ __device__ inline SlabPlan plan_slab(const Grid& G, int R, int Q, int col, int zz) {
    SlabPlan T;
    const int q0 = col * Q;
    T.T0 = (long long)zz * G.S2 + (long long)q0 * R * G.S;
    const int left = G.n2 - q0 * R;
    T.lines = left < Q * R ? left : Q * R;
    const long long T1 = T.T0 + (long long)T.lines * G.S, nev = (G.n + 1) & ~1LL, n8 = (G.n + 7) & ~7LL;
    if (zz < 0 || zz >= G.n3) { T.xa = T.xe = T.ba = T.be = T.pa = T.pe = 0; return T; }
    long long a = T.T0 - G.S - 1, e = T1 + G.S + 1;
    a = a < 0 ? 0 : a;
    e = e > nev ? nev : e;
    T.xa = a & ~1LL;
    T.xe = (e + 1) & ~1LL;
    T.ba = T.T0 & ~1LL;
    T.be = ((T1 + 1) & ~1LL) > nev ? nev : ((T1 + 1) & ~1LL);
    T.pa = T.T0 & ~7LL;
    T.pe = ((T1 + 7) & ~7LL) > n8 ? n8 : ((T1 + 7) & ~7LL);
    return T;
}
// the view of plane z of a column from the ring: slot(zz) = zz & 3; slot layout [x: xcap][b: bcap][pid: pcap]
__host__ __device__ inline View march_view(const Grid& G, int R, int Q, int col, int z, const unsigned char* ring,
                                           size_t slot_bytes, int xcap, int bcap) {
    View V;
    const SlabPlan C = plan_slab(G, R, Q, col, z);
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const int zz = z + d - 1;
        const SlabPlan P = plan_slab(G, R, Q, col, zz);
        const double* sx = reinterpret_cast<const double*>(ring + (size_t)(zz & 3) * slot_bytes);
        V.xs[d] = sx + (P.T0 - P.xa);   // xs[d][rel] = slab_zz[(T0(zz) + rel) - xa(zz)], T0(zz) = T0(z) + (d-1)*S2
    }
    const double* sc = reinterpret_cast<const double*>(ring + (size_t)(z & 3) * slot_bytes);
    V.b = sc + xcap + (C.T0 - C.ba);
    V.pid = reinterpret_cast<const uint16_t*>(sc + xcap + bcap) + (C.T0 - C.pa);
    return V;
}

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    const uint32_t a = smem_u32(bar);
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    } while (!ok);
}

// persistent CTAs of ceil32(Q*S) threads, two stages; x, b, pid need 16-byte aligned bases and (for x) an allocation that
// covers the even-rounded end of the vector
template <int R, int NTMAX>
__global__ void __launch_bounds__(NTMAX) k_lines_tma(Grid G, int Q, long long ntiles, const uint16_t* __restrict__ pid,
                                                    const Pat* __restrict__ pat, const Ent* __restrict__ ent,
                                                    const double* __restrict__ dp, const double* __restrict__ x,
                                                    const double* __restrict__ b, double* __restrict__ y) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
    int xcap, bcap, pcap;
    stage_layout(G, R, Q, xcap, bcap, pcap);
    const size_t stage_bytes = (((size_t)(3 * xcap + bcap) * 8 + (size_t)pcap * 2) + 127) / 128 * 128;
    unsigned char* stage0 = smem_raw + 128;
    const int t = threadIdx.x;
    if (t == 0) {
        mbar_init(full, 1);
        mbar_init(full + 1, 1);
        fence_mbar_init();
    }
    __syncthreads();
    auto issue = [&](long long tile, int s) {
        const TilePlan T = plan_tile(G, R, Q, tile);
        double* sx = reinterpret_cast<double*>(stage0 + (size_t)s * stage_bytes);
        double* sb = sx + 3 * (size_t)xcap;
        uint16_t* sp = reinterpret_cast<uint16_t*>(sb + bcap);
        uint32_t bytes = (uint32_t)(T.be - T.ba) * 8u + (uint32_t)(T.pe - T.pa) * 2u;
#pragma unroll
        for (int d = 0; d < 3; ++d)
            if (T.xe[d] > T.xa[d]) bytes += (uint32_t)(T.xe[d] - T.xa[d]) * 8u;
        mbar_expect_tx(full + s, bytes);
#pragma unroll
        for (int d = 0; d < 3; ++d)
            if (T.xe[d] > T.xa[d]) bulk_g2s(sx + (size_t)d * xcap, x + T.xa[d], (uint32_t)(T.xe[d] - T.xa[d]) * 8u, full + s);
        bulk_g2s(sb, b + T.ba, (uint32_t)(T.be - T.ba) * 8u, full + s);
        bulk_g2s(sp, pid + T.pa, (uint32_t)(T.pe - T.pa) * 2u, full + s);
    };
    if (t == 0 && blockIdx.x < ntiles) issue(blockIdx.x, 0);
    const int grp = t / G.S, i = t - grp * G.S;
    int it = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int s = it & 1;
        if (t == 0 && tile + gridDim.x < ntiles) issue(tile + gridDim.x, s ^ 1);
        mbar_wait(full + s, (it >> 1) & 1);
        const TilePlan T = plan_tile(G, R, Q, tile);
        const double* sx = reinterpret_cast<const double*>(stage0 + (size_t)s * stage_bytes);
        const double* sb = sx + 3 * (size_t)xcap;
        const uint16_t* sp = reinterpret_cast<const uint16_t*>(sb + bcap);
        if (grp < Q && grp * R < T.lines) {
            const View V = stage_view(G, T, sx, xcap, sb, sp);
            const long long rel0 = (long long)grp * R * G.S + i;
            const int y0 = (int)((T.T0 % G.S2) / G.S) + grp * R;
            lines_thread<R>(G, V, rel0, T.T0 + rel0, y0, pat, ent, dp, y);
        }
        __syncthreads();
    }
}
// work item = (column, z-chunk); ceil32(Q*S) threads
template <int R, int NTMAX>
__global__ void __launch_bounds__(NTMAX) k_lines_march(Grid G, int Q, int ncols, int zc, int nitems, const uint16_t* __restrict__ pid,
                                                       const Pat* __restrict__ pat, const Ent* __restrict__ ent,
                                                       const double* __restrict__ dp, const double* __restrict__ x,
                                                       const double* __restrict__ b, double* __restrict__ y) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
    int xcap, bcap, pcap;
    stage_layout(G, R, Q, xcap, bcap, pcap);
    const size_t slot_bytes = (((size_t)(xcap + bcap) * 8 + (size_t)pcap * 2) + 127) / 128 * 128;
    unsigned char* ring = smem_raw + 128;
    const int t = threadIdx.x;
    if (t == 0) {
        for (int k = 0; k < 4; ++k) mbar_init(full + k, 1);
        fence_mbar_init();
    }
    __syncthreads();
    uint32_t phase = 0;                                            // bit k: parity the next wait on slot k expects
    const int grp = t / G.S, i = t - grp * G.S;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        const int col = item % ncols, chunk = item / ncols;
        const int z0 = chunk * zc, z1 = min(z0 + zc, G.n3);
        const int first = max(z0 - 1, 0), last = min(z1, G.n3 - 1);
        auto issue = [&](int zz) {
            const SlabPlan P = plan_slab(G, R, Q, col, zz);
            const int k = zz & 3;
            double* sx = reinterpret_cast<double*>(ring + (size_t)k * slot_bytes);
            const bool own = (zz >= z0 && zz < z1);                // planes this item computes also need b and pid
            uint32_t bytes = (uint32_t)(P.xe - P.xa) * 8u;
            if (own) bytes += (uint32_t)(P.be - P.ba) * 8u + (uint32_t)(P.pe - P.pa) * 2u;
            mbar_expect_tx(full + k, bytes);
            bulk_g2s(sx, x + P.xa, (uint32_t)(P.xe - P.xa) * 8u, full + k);
            if (own) {
                bulk_g2s(sx + xcap, b + P.ba, (uint32_t)(P.be - P.ba) * 8u, full + k);
                bulk_g2s(reinterpret_cast<uint16_t*>(sx + xcap + bcap), pid + P.pa, (uint32_t)(P.pe - P.pa) * 2u, full + k);
            }
        };
        int issued = first - 1, waited = first - 1;
        if (t == 0)
            for (int zz = first; zz <= min(first + 2, last); ++zz) issue(zz);
        issued = min(first + 2, last);
        for (int z = z0; z < z1; ++z) {
            if (z + 2 <= last && z + 2 > issued) {                  // slot (z+2)&3 held plane z-2: free since the last barrier
                if (t == 0) issue(z + 2);
                issued = z + 2;
            }
            const int need = min(z + 1, last);
            while (waited < need) {
                ++waited;
                const int k = waited & 3;
                mbar_wait(full + k, (phase >> k) & 1);
                phase ^= 1u << k;
            }
            const SlabPlan C = plan_slab(G, R, Q, col, z);
            if (grp < Q && grp * R < C.lines) {
                const View V = march_view(G, R, Q, col, z, ring, slot_bytes, xcap, bcap);
                const long long rel0 = (long long)grp * R * G.S + i;
                lines_thread<R>(G, V, rel0, C.T0 + rel0, col * Q * R + grp * R, pat, ent, dp, y);
            }
            __syncthreads();
        }
    }
}
#endif

// pure streaming floor with the same vector traffic
__global__ void __launch_bounds__(256) k_stream(long long n, const uint16_t* __restrict__ pid, const double* __restrict__ dp,
                                                const double* __restrict__ x, const double* __restrict__ b, double* __restrict__ y) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) y[i] = x[i] + dp[pid[i]] * (b[i] - x[i]);
}

// synthetic dictionary matrix of an n1 x n2 x n3 nodal grid: 27 patterns (first / interior / last per dimension)
static void build(const Grid& G, int pts, std::vector<uint16_t>& pid, std::vector<Pat>& pat, std::vector<Ent>& ent,
                  std::vector<double>& dp) {
    pid.resize(G.n);
    pat.clear(); ent.clear(); dp.clear();
    for (int c = 0; c < 27; ++c) {
        const int cx = c % 3, cy = (c / 3) % 3, cz = c / 9;
        Pat P; P.k0 = (int)ent.size(); P.mask = 0;
        for (int dz = -1; dz <= 1; ++dz)
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                    const int nz = (dx != 0) + (dy != 0) + (dz != 0);
                    if (pts == 7 && nz > 1) continue;
                    if ((cx == 0 && dx < 0) || (cx == 2 && dx > 0) || (cy == 0 && dy < 0) || (cy == 2 && dy > 0) ||
                        (cz == 0 && dz < 0) || (cz == 2 && dz > 0)) continue;
                    Ent e;
                    e.v = nz == 0 ? 6.0 + 0.01 * c : -1.0 / (1 + nz) - 0.001 * c - 0.0001 * (dx + 3 * dy + 9 * dz);
                    e.delta = dx + G.S * dy + G.S2 * dz;
                    e.pad = 0;
                    ent.push_back(e);
                    P.mask |= 1 << ((dz + 1) * 9 + (dy + 1) * 3 + (dx + 1));
                }
        P.len = (int)ent.size() - P.k0;
        pat.push_back(P);
        dp.push_back(0.8 / (6.0 + 0.01 * c));
    }
    for (int k = 0; k < G.n3; ++k)
        for (int j = 0; j < G.n2; ++j)
            for (int i = 0; i < G.n1; ++i) {
                const int cx = i == 0 ? 0 : (i == G.n1 - 1 ? 2 : 1), cy = j == 0 ? 0 : (j == G.n2 - 1 ? 2 : 1),
                          cz = k == 0 ? 0 : (k == G.n3 - 1 ? 2 : 1);
                pid[(size_t)k * G.S2 + (size_t)j * G.S + i] = (uint16_t)(cx + 3 * cy + 9 * cz);
            }
}

static Grid make_grid(int n1, int n2, int n3) {
    Grid G; G.n1 = n1; G.n2 = n2; G.n3 = n3; G.S = n1; G.S2 = n1 * n2; G.n = (long long)n1 * n2 * n3;
    return G;
}

template <int R>
static bool host_check_one(const Grid& G, int pts) {
    std::vector<uint16_t> pid; std::vector<Pat> pat; std::vector<Ent> ent; std::vector<double> dp;
    build(G, pts, pid, pat, ent, dp);
    pid.resize(G.n + 16, 0xFFFF);
    const long long pad = G.S2 + G.S + 1;
    std::vector<double> xbuf(G.n + 2 * pad, 1e300), b(G.n + 2, 1e300), yref(G.n), y(G.n, -1.0), y2(G.n, -1.0);   // poison in the padding
    double* x = xbuf.data() + pad;
    srand(7);
    for (long long i = 0; i < G.n; ++i) { x[i] = rand() / (double)RAND_MAX; b[i] = rand() / (double)RAND_MAX; }
    for (long long r = 0; r < G.n; ++r) yref[r] = row_reference(G, r, pid.data(), pat.data(), ent.data(), dp.data(), x, b.data());
    // (a) global-memory form
    const int gpp = (G.n2 + R - 1) / R;
    const View Vg = global_view(G, pid.data(), x, b.data());
    for (int z = 0; z < G.n3; ++z)
        for (int q = 0; q < gpp; ++q)
            for (int i = 0; i < G.n1; ++i) {
                const long long row0 = (long long)z * G.S2 + (long long)q * R * G.S + i;
                lines_thread<R>(G, Vg, row0, row0, q * R, pat.data(), ent.data(), dp.data(), y.data());
            }
    bool ok = memcmp(y.data(), yref.data(), G.n * sizeof(double)) == 0;
    // (b) staged form: the stage is a host buffer filled exactly as the kernel's bulk copies fill shared memory
    for (int Q = 1; Q <= 3; ++Q) {
        std::fill(y2.begin(), y2.end(), -1.0);
        int xcap, bcap, pcap;
        stage_layout(G, R, Q, xcap, bcap, pcap);
        const int tpp = (gpp + Q - 1) / Q;
        const long long ntiles = (long long)G.n3 * tpp;
        std::vector<double> sx(3 * (size_t)xcap), sb(bcap);
        std::vector<uint16_t> sp(pcap);
        for (long long tile = 0; tile < ntiles; ++tile) {
            const TilePlan T = plan_tile(G, R, Q, tile);
            std::fill(sx.begin(), sx.end(), 1e300);
            std::fill(sb.begin(), sb.end(), 1e300);
            std::fill(sp.begin(), sp.end(), (uint16_t)0xFFFF);
            for (int d = 0; d < 3; ++d) {
                if (T.xe[d] - T.xa[d] > xcap) { printf("x range exceeds the stage\n"); return false; }
                for (long long g = T.xa[d]; g < T.xe[d]; ++g) sx[(size_t)d * xcap + (g - T.xa[d])] = x[g];   // x[n] (even rounding) is padding
            }
            if (T.be - T.ba > bcap || T.pe - T.pa > pcap) { printf("b / pid range exceeds the stage\n"); return false; }
            for (long long g = T.ba; g < T.be; ++g) sb[g - T.ba] = b[g];
            for (long long g = T.pa; g < T.pe; ++g) sp[g - T.pa] = pid[g];
            const View V = stage_view(G, T, sx.data(), xcap, sb.data(), sp.data());
            const int nthreads = ((Q * G.S + 31) / 32) * 32;
            for (int t = 0; t < nthreads; ++t) {
                const int grp = t / G.S, i = t - grp * G.S;
                if (!(grp < Q && grp * R < T.lines)) continue;
                const long long rel0 = (long long)grp * R * G.S + i;
                const int y0 = (int)((T.T0 % G.S2) / G.S) + grp * R;
                lines_thread<R>(G, V, rel0, T.T0 + rel0, y0, pat.data(), ent.data(), dp.data(), y2.data());
            }
        }
        ok = ok && memcmp(y2.data(), yref.data(), G.n * sizeof(double)) == 0;
    }
    // (c) marching form: the ring of four slots is replayed literally - a slot is poisoned and refilled exactly when the
    // kernel would issue its copies, so a wrong schedule (a slab overwritten while still needed) shows as a mismatch
    for (int Q = 1; Q <= 2; ++Q)
        for (int zc : {1, 3, G.n3}) {
            std::fill(y2.begin(), y2.end(), -1.0);
            int xcap, bcap, pcap;
            stage_layout(G, R, Q, xcap, bcap, pcap);
            const size_t slot_bytes = (((size_t)(xcap + bcap) * 8 + (size_t)pcap * 2) + 127) / 128 * 128;
            std::vector<unsigned char> ringv(4 * slot_bytes + 16);
            unsigned char* ring = ringv.data() + (16 - ((uintptr_t)ringv.data() & 15)) % 16;
            const int ncols = (gpp + Q - 1) / Q, nchunks = (G.n3 + zc - 1) / zc;
            for (int item = 0; item < ncols * nchunks; ++item) {
                const int col = item % ncols, chunk = item / ncols;
                const int z0 = chunk * zc, z1 = std::min(z0 + zc, G.n3);
                const int first = std::max(z0 - 1, 0), last = std::min(z1, G.n3 - 1);
                auto issue = [&](int zz) {
                    const SlabPlan P = plan_slab(G, R, Q, col, zz);
                    double* sx = reinterpret_cast<double*>(ring + (size_t)(zz & 3) * slot_bytes);
                    for (size_t k = 0; k < slot_bytes / 8; ++k) sx[k] = 1e300;
                    if (P.xe - P.xa > xcap || P.be - P.ba > bcap || P.pe - P.pa > pcap) { printf("slab exceeds the slot\n"); exit(1); }
                    for (long long g = P.xa; g < P.xe; ++g) sx[g - P.xa] = x[g];
                    if (zz >= z0 && zz < z1) {
                        for (long long g = P.ba; g < P.be; ++g) sx[xcap + (g - P.ba)] = b[g];
                        uint16_t* sp = reinterpret_cast<uint16_t*>(sx + xcap + bcap);
                        for (long long g = P.pa; g < P.pe; ++g) sp[g - P.pa] = pid[g];
                    }
                };
                int issued = std::min(first + 2, last);
                for (int zz = first; zz <= issued; ++zz) issue(zz);
                for (int z = z0; z < z1; ++z) {
                    if (z + 2 <= last && z + 2 > issued) { issue(z + 2); issued = z + 2; }
                    const SlabPlan C = plan_slab(G, R, Q, col, z);
                    const int nthreads = ((Q * G.S + 31) / 32) * 32;
                    for (int t = 0; t < nthreads; ++t) {
                        const int grp = t / G.S, i = t - grp * G.S;
                        if (!(grp < Q && grp * R < C.lines)) continue;
                        const View V = march_view(G, R, Q, col, z, ring, slot_bytes, xcap, bcap);
                        const long long rel0 = (long long)grp * R * G.S + i;
                        lines_thread<R>(G, V, rel0, C.T0 + rel0, col * Q * R + grp * R, pat.data(), ent.data(), dp.data(), y2.data());
                    }
                }
            }
            ok = ok && memcmp(y2.data(), yref.data(), G.n * sizeof(double)) == 0;
        }
    printf("host check %2d-point %dx%dx%d R=%d (global form, staged form Q=1..3, marching form): %s\n", pts, G.n1, G.n2, G.n3, R,
           ok ? "bit-identical" : "MISMATCH");
    return ok;
}

static int host_check() {
    bool ok = true;
    const int grids[4][3] = {{9, 9, 9}, {12, 9, 7}, {5, 11, 3}, {33, 6, 4}};
    for (auto& g : grids)
        for (int pts : {7, 27}) {
            const Grid G = make_grid(g[0], g[1], g[2]);
            ok = host_check_one<2>(G, pts) && ok;
            ok = host_check_one<4>(G, pts) && ok;
            ok = host_check_one<8>(G, pts) && ok;
        }
    return ok ? 0 : 1;
}

int main(int argc, char** argv) {
    if (argc > 1 && strcmp(argv[1], "--host-check") == 0) return host_check();
    const int reps = 20;
    int nsm = 0;
    CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
    for (int cfg = 0; cfg < 2; ++cfg) {
        const int N = cfg == 0 ? 257 : 129, pts = cfg == 0 ? 7 : 27;
        const Grid G = make_grid(N, N, N);
        const long long n = G.n, pad = (G.S2 + G.S + 2) & ~1LL;   // even: x[0] stays 16-byte aligned
        std::vector<uint16_t> pid; std::vector<Pat> pat; std::vector<Ent> ent; std::vector<double> dp;
        build(G, pts, pid, pat, ent, dp);
        std::vector<double> hx(n), hb(n), href(n), hy(n);
        srand(1);
        for (long long i = 0; i < n; ++i) { hx[i] = rand() / (double)RAND_MAX; hb[i] = rand() / (double)RAND_MAX; }
        uint16_t* dpid; Pat* dpat; Ent* dent; double *ddp, *dxb, *db, *dy;
        CK(cudaMalloc(&dpid, n * 2 + 64)); CK(cudaMalloc(&dpat, pat.size() * sizeof(Pat))); CK(cudaMalloc(&dent, ent.size() * sizeof(Ent)));
        CK(cudaMalloc(&ddp, dp.size() * 8)); CK(cudaMalloc(&dxb, (n + 2 * pad) * 8)); CK(cudaMalloc(&db, n * 8)); CK(cudaMalloc(&dy, n * 8));
        CK(cudaMemset(dxb, 0, (n + 2 * pad) * 8));
        double* dx = dxb + pad;
        CK(cudaMemcpy(dpid, pid.data(), n * 2, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dpat, pat.data(), pat.size() * sizeof(Pat), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dent, ent.data(), ent.size() * sizeof(Ent), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(ddp, dp.data(), dp.size() * 8, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dx, hx.data(), n * 8, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(db, hb.data(), n * 8, cudaMemcpyHostToDevice));
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        const int grid1 = (int)((n + 255) / 256);
        const char* names[] = {"reference 1 row/thread", "stream floor", "lines R=2 x4 CTAs/SM", "lines R=2 x8 CTAs/SM", "lines R=4 x3 CTAs/SM",
                               "lines R=4 x6 CTAs/SM", "lines R=8 x2 CTAs/SM", "lines R=8 x4 CTAs/SM", "lines TMA R=2 Q=1", "lines TMA R=2 Q=2",
                               "lines TMA R=4 Q=1", "lines TMA R=4 Q=2", "lines TMA R=8 Q=1", "march R=2 Q=1", "march R=4 Q=1", "march R=4 Q=2",
                               "march R=8 Q=1"};
        // staged form: Q groups per tile, ceil32(Q*S) threads, 2 stages; as many CTAs per SM as shared memory allows
        auto tma_launch = [&](int R, int Q) {
            int xcap, bcap, pcap;
            stage_layout(G, R, Q, xcap, bcap, pcap);
            const size_t stage = (((size_t)(3 * xcap + bcap) * 8 + (size_t)pcap * 2) + 127) / 128 * 128, smem = 128 + 2 * stage;
            const int nt = ((Q * G.S + 31) / 32) * 32;
            const int gpp = (G.n2 + R - 1) / R, tpp = (gpp + Q - 1) / Q;
            const long long ntiles = (long long)G.n3 * tpp;
            if (smem > 227 * 1024 || nt > 544) { printf("  (R=%d Q=%d skipped: %zu B shared memory, %d threads)\n", R, Q, smem, nt); return; }
            int per = (int)std::min<size_t>((size_t)(2048 / nt), (228 * 1024) / (smem + 1024));
            if (per < 1) per = 1;
            const int grid = (int)std::min<long long>(ntiles, (long long)nsm * per);
#define TL(R_) { static bool once = false; if (!once) { CK(cudaFuncSetAttribute(k_lines_tma<R_, 544>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); once = true; } \
                 k_lines_tma<R_, 544><<<grid, nt, smem>>>(G, Q, ntiles, dpid, dpat, dent, ddp, dx, db, dy); }
            if (R == 2) TL(2) else if (R == 4) TL(4) else TL(8)
#undef TL
        };
        // marching form: items = columns x z-chunks, chunks sized so that about 3 items per resident CTA exist
        auto march_launch = [&](int R, int Q) {
            int xcap, bcap, pcap;
            stage_layout(G, R, Q, xcap, bcap, pcap);
            const size_t slot = (((size_t)(xcap + bcap) * 8 + (size_t)pcap * 2) + 127) / 128 * 128, smem = 128 + 4 * slot;
            const int nt = ((Q * G.S + 31) / 32) * 32;
            if (smem > 227 * 1024 || nt > 544) { printf("  (march R=%d Q=%d skipped: %zu B shared memory, %d threads)\n", R, Q, smem, nt); return; }
            int per = (int)std::min<size_t>((size_t)(2048 / nt), (228 * 1024) / (smem + 1024));
            if (per < 1) per = 1;
            const int gpp = (G.n2 + R - 1) / R, ncols = (gpp + Q - 1) / Q;
            int nchunks = std::max(1, (3 * nsm * per + ncols - 1) / ncols);
            int zc = std::max(8, (G.n3 + nchunks - 1) / nchunks);
            nchunks = (G.n3 + zc - 1) / zc;
            const int nitems = ncols * nchunks, grid = std::min(nitems, nsm * per);
#define ML(R_) { static bool once = false; if (!once) { CK(cudaFuncSetAttribute(k_lines_march<R_, 544>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); once = true; } \
                 k_lines_march<R_, 544><<<grid, nt, smem>>>(G, Q, ncols, zc, nitems, dpid, dpat, dent, ddp, dx, db, dy); }
            if (R == 2) ML(2) else if (R == 4) ML(4) else ML(8)
#undef ML
        };
        for (int v = 0; v < 17; ++v) {
            auto launch = [&]() {
                switch (v) {
                    case 0: k_reference<<<grid1, 256>>>(G, dpid, dpat, dent, ddp, dx, db, dy); break;
                    case 1: k_stream<<<grid1, 256>>>(n, dpid, ddp, dx, db, dy); break;
                    case 2: k_lines<2><<<nsm * 4, 256>>>(G, dpid, dpat, dent, ddp, dx, db, dy); break;
                    case 3: k_lines<2><<<nsm * 8, 256>>>(G, dpid, dpat, dent, ddp, dx, db, dy); break;
                    case 4: k_lines<4><<<nsm * 3, 256>>>(G, dpid, dpat, dent, ddp, dx, db, dy); break;
                    case 5: k_lines<4><<<nsm * 6, 256>>>(G, dpid, dpat, dent, ddp, dx, db, dy); break;
                    case 6: k_lines<8><<<nsm * 2, 256>>>(G, dpid, dpat, dent, ddp, dx, db, dy); break;
                    case 7: k_lines<8><<<nsm * 4, 256>>>(G, dpid, dpat, dent, ddp, dx, db, dy); break;
                    case 8: tma_launch(2, 1); break;
                    case 9: tma_launch(2, 2); break;
                    case 10: tma_launch(4, 1); break;
                    case 11: tma_launch(4, 2); break;
                    case 12: tma_launch(8, 1); break;
                    case 13: march_launch(2, 1); break;
                    case 14: march_launch(4, 1); break;
                    case 15: march_launch(4, 2); break;
                    case 16: march_launch(8, 1); break;
                }
            };
            CK(cudaMemset(dy, 0, n * 8));
            for (int w = 0; w < 3; ++w) launch();
            CK(cudaGetLastError());
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0));
            for (int r = 0; r < reps; ++r) launch();
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            CK(cudaMemcpy(hy.data(), dy, n * 8, cudaMemcpyDeviceToHost));
            bool same = true;
            if (v == 0) href = hy;
            else if (v != 1) same = memcmp(href.data(), hy.data(), n * 8) == 0;
            const double us = 1e3 * ms / reps;
            printf("%2d-point N=%d  %-24s %8.1f us  %7.1f GB/s (26 B/row)  %s\n", pts, N, names[v], us, 26.0 * n / us / 1e3,
                   v != 1 ? (same ? "bit-identical" : "MISMATCH") : "");
        }
        cudaFree(dpid); cudaFree(dpat); cudaFree(dent); cudaFree(ddp); cudaFree(dxb); cudaFree(db); cudaFree(dy);
    }
    return 0;
}
