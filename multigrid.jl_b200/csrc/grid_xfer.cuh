// Transfer operators of the geometric hierarchy with a GRID HINT (option "grid_transfers", on by default).
//
// P = kron of 1-D linear interpolations on nodal grids, R = 2^-dim P^T (src/Multigrid/GeometricTransferOperators.jl:5-46,
// MGsetup.jl:53-62 of the reference).  In stencil-dictionary form (pattern.cuh) their rows are copies of 8 (P: one per
// parity class of the fine node) or 27 (R: first / interior / last per dimension) patterns, but the dictionary kernel
// still reads 6 bytes per row (pattern id + first column) and walks the entries per lane with 1 ... 8 trips inside a
// warp (prolongation) or with a 16-byte entry load per product (restriction).  The reference's setup knows the meshes
// (param.Meshes[l].n); when the host passes them (mgb200_set_level_grid) and the uploaded P / R are exactly what those
// grids imply - verified row by row at upload, values untouched - the pattern and the columns of a row are functions
// of its grid coordinates.  The kernels below then read NO matrix stream at all, only dense tables of the dictionary
// VALUES (so a P with other weights still works):
//   prolongation: a CTA walks fine lines; a thread owns fine node i of the line (coalesced read-modify-write of x_f); the
//                 parities of the line (b, c) are CTA-uniform, the products of a row run in stored order (dz', dy', dx');
//   restriction:  a CTA walks coarse lines; a thread owns coarse node I; 27 predicated products in stored order with the
//                 coefficients of the line's boundary class in shared memory.
// One thread = one row and stored order, so results are bit-identical to the dictionary walk.  The per-row functions
// are __host__ __device__: mgb200_host_grid_transfer runs them on the CPU (tests/test_patterns.py).
// (Round 1 had a coarse-column-per-thread form with R lines per thread; measured on the B200 it gained little -
// profiles/r02a_microbench_lines_transfer.log - and was replaced by the kernels below.)
#pragma once
#include "pattern.cuh"
#include "ll.cuh"

namespace mgb200 {

struct GridXfer {
    int ok;            // 0: no hint / hint does not match the matrix
    int kind;          // 1: prolongation (rows = fine nodes), 2: restriction (rows = coarse nodes)
    int n[3], N[3];    // fine / coarse nodes per dimension (unused dimensions: 1); n = 2N - 1 where N > 1
    int cls_k0[27];    // first dictionary entry of the pattern of each class (-1: class does not occur)
    void* tab;         // device: dense value table (gx_dense_table), released with the matrix
    // z-slab of a row-partitioned level: the matrix holds the rows of the planes [k0, k0 + nk) of the ROW grid (fine
    // planes for P, coarse planes for R); its column indices are global indices of the column grid minus `shift`
    int k0, nk;
    long long shift;
};
constexpr int GXP_TAB = 8 * 8;      // prolongation: class a + 2b + 4c, up to 8 values in stored order (zeros beyond)
constexpr int GXR_TAB = 27 * 27;    // restriction: class cx + 3cy + 9cz, coefficient of offset (dz+1)*9 + (dy+1)*3 + (dx+1)
static inline GridXfer no_grid() {
    GridXfer X;
    std::memset(&X, 0, sizeof(X));
    return X;
}

// offsets {-1,0,+1} (bit 0,1,2) a coarse node of boundary class c may use in a dimension with N coarse nodes
__host__ __device__ __forceinline__ int gx_allowed(int c, int N) { return N == 1 ? 2 : (c == 0 ? 6 : (c == 2 ? 3 : 7)); }
__host__ __device__ __forceinline__ int gx_class(int I, int N) { return I == 0 ? 0 : (I == N - 1 ? 2 : 1); }

// ---- upload-time verification (host) ---------------------------------------------------------------------------------
template <typename TA>
static bool gx_verify_prolongation(const HostPatterns<TA>& H, long long n_rows, const int n[3], const int N[3], GridXfer& X,
                                   int k0 = 0, int nk = -1, long long shift = 0) {
    X = no_grid();
    if (!H.ok || H.rowrel) return false;
    for (int d = 0; d < 3; ++d) {
        if (N[d] < 1 || n[d] != (N[d] > 1 ? 2 * N[d] - 1 : 1)) return false;
        X.n[d] = n[d];
        X.N[d] = N[d];
    }
    if (nk < 0) nk = n[2];
    if (k0 < 0 || k0 + nk > n[2] || (long long)n[0] * n[1] * nk != n_rows) return false;
    X.k0 = k0;
    X.nk = nk;
    X.shift = shift;
    int cls_pat[8];
    for (int c = 0; c < 8; ++c) cls_pat[c] = -1;
    for (int c = 0; c < 27; ++c) X.cls_k0[c] = -1;
    const long long N1 = N[0], N12 = (long long)N[0] * N[1];
    long long row = 0;
    for (int k = k0; k < k0 + nk; ++k)
        for (int j = 0; j < n[1]; ++j)
            for (int i = 0; i < n[0]; ++i, ++row) {
                const int a = i & 1, b = j & 1, c = k & 1, cls = a + 2 * b + 4 * c, p = H.pid[row];
                if (H.c0[row] != (i >> 1) + N1 * (j >> 1) + N12 * (k >> 1) - shift) return false;
                if (cls_pat[cls] == p) continue;
                if (cls_pat[cls] != -1) return false;
                int q = H.pat_off[p];
                if (H.pat_off[p + 1] - q != (1 << (a + b + c))) return false;
                for (int dz = 0; dz <= c; ++dz)
                    for (int dy = 0; dy <= b; ++dy)
                        for (int dx = 0; dx <= a; ++dx, ++q)
                            if (H.delta[q] != dx + N1 * dy + N12 * dz) return false;
                cls_pat[cls] = p;
                X.cls_k0[cls] = H.pat_off[p];
            }
    X.kind = 1;
    X.ok = 1;
    return true;
}
template <typename TA>
static bool gx_verify_restriction(const HostPatterns<TA>& H, long long n_rows, const int n[3], const int N[3], GridXfer& X,
                                  int k0 = 0, int nk = -1, long long shift = 0) {
    X = no_grid();
    if (!H.ok || H.rowrel) return false;
    for (int d = 0; d < 3; ++d) {
        if (N[d] < 1 || n[d] != (N[d] > 1 ? 2 * N[d] - 1 : 1)) return false;
        X.n[d] = n[d];
        X.N[d] = N[d];
    }
    if (nk < 0) nk = N[2];
    if (k0 < 0 || k0 + nk > N[2] || (long long)N[0] * N[1] * nk != n_rows) return false;
    X.k0 = k0;
    X.nk = nk;
    X.shift = shift;
    int cls_pat[27];
    for (int c = 0; c < 27; ++c) {
        cls_pat[c] = -1;
        X.cls_k0[c] = -1;
    }
    const long long S = n[0], S2 = (long long)n[0] * n[1];
    long long row = 0;
    for (int K = k0; K < k0 + nk; ++K)
        for (int J = 0; J < N[1]; ++J)
            for (int I = 0; I < N[0]; ++I, ++row) {
                const int cx = gx_class(I, N[0]), cy = gx_class(J, N[1]), cz = gx_class(K, N[2]), cls = cx + 3 * cy + 9 * cz;
                const int ax = gx_allowed(cx, N[0]), ay = gx_allowed(cy, N[1]), az = gx_allowed(cz, N[2]);
                const int p = H.pid[row];
                const long long anchor = 2LL * I + S * (2LL * J) + S2 * (2LL * K);
                // first entry = smallest allowed offset in every dimension
                const int fx = (ax & 1) ? -1 : 0, fy = (ay & 1) ? -1 : 0, fz = (az & 1) ? -1 : 0;
                const long long first = fx + S * fy + S2 * fz;
                if (H.c0[row] != anchor + first - shift) return false;
                if (cls_pat[cls] == p) continue;
                if (cls_pat[cls] != -1) return false;
                int q = H.pat_off[p];
                const int q1 = H.pat_off[p + 1];
                for (int dz = -1; dz <= 1; ++dz)
                    for (int dy = -1; dy <= 1; ++dy)
                        for (int dx = -1; dx <= 1; ++dx) {
                            if (!((az >> (dz + 1)) & 1) || !((ay >> (dy + 1)) & 1) || !((ax >> (dx + 1)) & 1)) continue;
                            if (q >= q1 || H.delta[q] != (dx + S * dy + S2 * dz) - first) return false;
                            ++q;
                        }
                if (q != q1) return false;
                cls_pat[cls] = p;
                X.cls_k0[cls] = H.pat_off[p];
            }
    X.kind = 2;
    X.ok = 1;
    return true;
}

// dense value table of a verified hint, from the dictionary values (host)
template <typename TA>
static std::vector<TA> gx_dense_table(const GridXfer& X, const std::vector<TA>& val) {
    std::vector<TA> t(X.kind == 1 ? GXP_TAB : GXR_TAB, TA());
    if (X.kind == 1) {
        for (int cls = 0; cls < 8; ++cls) {
            if (X.cls_k0[cls] < 0) continue;
            const int a = cls & 1, b = (cls >> 1) & 1, c = (cls >> 2) & 1;
            for (int k = 0; k < (1 << (a + b + c)); ++k) t[cls * 8 + k] = val[X.cls_k0[cls] + k];
        }
    } else {
        for (int cls = 0; cls < 27; ++cls) {
            if (X.cls_k0[cls] < 0) continue;
            const int ax = gx_allowed(cls % 3, X.N[0]), ay = gx_allowed((cls / 3) % 3, X.N[1]), az = gx_allowed(cls / 9, X.N[2]);
            int q = X.cls_k0[cls];
            for (int dz = -1; dz <= 1; ++dz)
                for (int dy = -1; dy <= 1; ++dy)
                    for (int dx = -1; dx <= 1; ++dx)
                        if (((az >> (dz + 1)) & 1) && ((ay >> (dy + 1)) & 1) && ((ax >> (dx + 1)) & 1))
                            t[cls * 27 + (dz + 1) * 9 + (dy + 1) * 3 + (dx + 1)] = val[q++];
        }
    }
    return t;
}

template <typename T>
__host__ __device__ __forceinline__ T gx_ld(const T* p) {
#ifdef __CUDA_ARCH__
    return ldg_(p);
#else
    return *p;
#endif
}

// ---- prolongation: x_f[i] += sum of P's row, fine node (i, j, k); b = j & 1, c = k & 1; q = coarse node (0, j/2, k/2) ----
// tab: the 8 x 8 dense table.  Products in stored order (dz', dy', dx').
template <typename TA, typename TV>
__host__ __device__ __forceinline__ TV gxp_row(const TA* tab, const TV* q, int N0, long long cs2, int i, int b, int c, TV xf) {
    const int a = i & 1;
    const TA* tv = tab + (a + 2 * b + 4 * c) * 8;
    const TV* p0 = q + (i >> 1);
    TV acc = VT<TV>::zero();
    int idx = 0;
#pragma unroll
    for (int dz = 0; dz < 2; ++dz) {
        if (dz > c) break;
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) {
            if (dy > b) break;
            const TV* p = p0 + dy * N0 + dz * cs2;
            acc = acc + tv[idx] * gx_ld(p);
            ++idx;
            if (a) {
                acc = acc + tv[idx] * gx_ld(p + 1);
                ++idx;
            }
        }
    }
    return xf + acc;
}
// ---- restriction: coarse node (I, J, K) = 27 predicated products around fine node (2I, 2J, 2K) in stored order -------------
// tv: the 27 coefficients of the node's class; ax / ay / az: allowed offsets per dimension (gx_allowed)
template <typename TA, typename TV>
__host__ __device__ __forceinline__ TV gxr_row(const TA* tv, const TV* f, long long S, long long S2, int ax, int ay, int az) {
    TV acc = VT<TV>::zero();
#pragma unroll
    for (int dz = -1; dz <= 1; ++dz) {
        if (!((az >> (dz + 1)) & 1)) continue;
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
            if (!((ay >> (dy + 1)) & 1)) continue;
            const TV* p = f + dz * S2 + dy * S;
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx)
                if ((ax >> (dx + 1)) & 1) acc = acc + tv[(dz + 1) * 9 + (dy + 1) * 3 + (dx + 1)] * gx_ld(p + dx);
        }
    }
    return acc;
}

// persistent CTAs over the fine lines (j, k); threads over i.  (Two or four lines per thread with the x_f reads issued
// up front measured slower on the B200: 107 / 107 against 99 us at 257^3, profiles/r02_tune_log.txt.)
template <typename TA, typename TV>
__global__ void __launch_bounds__(1024) gxp_kernel(const __grid_constant__ GridXfer X, const __grid_constant__ PutPlan pp,
                                                   const TA* __restrict__ tabg, const TV* __restrict__ xc, TV* __restrict__ xf) {
    __shared__ TA tab[GXP_TAB];
    for (int i = threadIdx.x; i < GXP_TAB; i += blockDim.x) tab[i] = tabg[i];
    __syncthreads();
    const int n0 = X.n[0], n1 = X.n[1], N0 = X.N[0], N1 = X.N[1];
    const long long cs2 = (long long)N0 * N1;
    const int nlines = n1 * X.nk;                         // the rows of this matrix: fine planes k0 .. k0 + nk - 1
    for (int line = blockIdx.x; line < nlines; line += gridDim.x) {
        const int kl = line / n1, j = line - kl * n1, k = X.k0 + kl;
        const TV* q = xc + (((long long)(k >> 1) * N1 + (j >> 1)) * N0 - X.shift);
        TV* xl = xf + (long long)line * n0;
        for (int i = threadIdx.x; i < n0; i += blockDim.x) {
            const TV v = gxp_row<TA, TV>(tab, q, N0, cs2, i, j & 1, k & 1, xl[i]);
            xl[i] = v;
            if (pp.on) ll_put_edge<TV>(pp, (long long)line * n0 + i, v);
        }
    }
}
// ---- prolongation, quad form (the default) ----------------------------------------------------------------------------------
// The line form above is latency-bound (profiles/r02q_ncu_gxp_summary.txt: 144 us at 257^3, DRAM 20 %, ~143 warp
// instructions per 32 rows, one x_f load in flight per thread, the coarse loads behind parity branches).  Here a thread
// owns fine column i of the (up to) FOUR fine lines (2J + b, 2K + c) of one coarse cell row (J, K): the 4 ... 8 coarse
// values those lines interpolate from and the four x_f values are loaded first, unconditionally (line and plane
// existence is uniform over the thread's y-row; only the `a` parity is per lane), and the index arithmetic is paid
// once per four rows.  Each row is still the sum of its products in stored order (dz', dy', dx') starting from zero and
// added to x_f last - gxp_row above - so the result does not change by a bit.  blockDim.y cell rows per CTA pass keep
// narrow (coarse) levels from running one-warp CTAs.
template <typename TA, typename TV>
__device__ __forceinline__ TV gxp_quad_row(const TA* tab, const TV (&v)[2][2][2], int a, int b, int c, TV xf) {
    // compact table of class a + 2b + 4c: entry ((dz (b + 1) + dy) << a) + dx
    const TA* tv = tab + (a + 2 * b + 4 * c) * 8;
    TV acc = VT<TV>::zero();
#pragma unroll
    for (int dz = 0; dz < 2; ++dz) {
        if (dz > c) break;
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) {
            if (dy > b) break;
            const int e = (dz * (b + 1) + dy) << a;
            acc = acc + tv[e] * v[dz][dy][0];
            if (a) acc = acc + tv[e + 1] * v[dz][dy][1];
        }
    }
    return xf + acc;
}
template <typename TA, typename TV, int MINB>
__global__ void __launch_bounds__(512, MINB) gxp_quad_kernel(const __grid_constant__ GridXfer X, const __grid_constant__ PutPlan pp,
                                                        const TA* __restrict__ tabg, const TV* __restrict__ xc,
                                                        TV* __restrict__ xf) {
    __shared__ TA tab[GXP_TAB];
    for (int i = threadIdx.y * blockDim.x + threadIdx.x; i < GXP_TAB; i += blockDim.x * blockDim.y) tab[i] = tabg[i];
    __syncthreads();
    const int n0 = X.n[0], n1 = X.n[1], N0 = X.N[0], N1 = X.N[1];
    const long long cs2 = (long long)N0 * N1, fs2 = (long long)n0 * n1;
    const int k_end = X.k0 + X.nk;                         // the rows of this matrix: fine planes k0 .. k_end - 1
    const int K_first = X.k0 >> 1, nK = ((k_end - 1) >> 1) - K_first + 1;
    const int ngroups = N1 * nK;
    for (int g = blockIdx.x * blockDim.y + threadIdx.y; g < ngroups; g += gridDim.x * blockDim.y) {
        const int Kl = g / N1, J = g - Kl * N1, K = K_first + Kl;
        const int j0 = 2 * J, kk = 2 * K;
        const bool eb = j0 + 1 < n1;                                     // the lines with b = 1 exist
        const bool e0 = kk >= X.k0, e1 = kk + 1 < k_end;                 // the planes with c = 0 / c = 1 are rows of this matrix
        const TV* q00 = xc + (((long long)K * N1 + J) * N0 - X.shift);
        const TV* q10 = q00 + N0;
        const TV* q01 = q00 + cs2;
        const TV* q11 = q01 + N0;
        const long long r00 = ((long long)(kk - X.k0) * n1 + j0) * n0;   // first row of line (b, c) = (0, 0); negative if !e0
        for (int i = threadIdx.x; i < n0; i += blockDim.x) {
            const int a = i & 1, I = i >> 1;
            TV v[2][2][2];
#pragma unroll
            for (int z = 0; z < 8; ++z) v[z >> 2][(z >> 1) & 1][z & 1] = VT<TV>::zero();
            TV f00 = VT<TV>::zero(), f10 = f00, f01 = f00, f11 = f00;
            // all loads first
            v[0][0][0] = ldg_(q00 + I);
            if (a) v[0][0][1] = ldg_(q00 + I + 1);
            if (eb) {
                v[0][1][0] = ldg_(q10 + I);
                if (a) v[0][1][1] = ldg_(q10 + I + 1);
            }
            if (e1) {
                v[1][0][0] = ldg_(q01 + I);
                if (a) v[1][0][1] = ldg_(q01 + I + 1);
                if (eb) {
                    v[1][1][0] = ldg_(q11 + I);
                    if (a) v[1][1][1] = ldg_(q11 + I + 1);
                }
            }
            const long long r = r00 + i;
            if (e0) {
                f00 = xf[r];
                if (eb) f10 = xf[r + n0];
            }
            if (e1) {
                f01 = xf[r + fs2];
                if (eb) f11 = xf[r + fs2 + n0];
            }
            if (e0) {
                const TV o = gxp_quad_row<TA, TV>(tab, v, a, 0, 0, f00);
                xf[r] = o;
                if (pp.on) ll_put_edge<TV>(pp, r, o);
                if (eb) {
                    const TV o1 = gxp_quad_row<TA, TV>(tab, v, a, 1, 0, f10);
                    xf[r + n0] = o1;
                    if (pp.on) ll_put_edge<TV>(pp, r + n0, o1);
                }
            }
            if (e1) {
                const TV o = gxp_quad_row<TA, TV>(tab, v, a, 0, 1, f01);
                xf[r + fs2] = o;
                if (pp.on) ll_put_edge<TV>(pp, r + fs2, o);
                if (eb) {
                    const TV o1 = gxp_quad_row<TA, TV>(tab, v, a, 1, 1, f11);
                    xf[r + fs2 + n0] = o1;
                    if (pp.on) ll_put_edge<TV>(pp, r + fs2 + n0, o1);
                }
            }
        }
    }
}

// persistent CTAs over the coarse lines (J, K); threads over I.
// (Measured and rejected, profiles/r02r_*: 16-byte loads of the aligned (even, odd) fine pair with all 18 loads of a row
// issued up front need 80 registers; at 160-thread CTAs that left 11 warps per SM and the kernel ran at 107 us against
// 55 us for this form at 257^3.)
template <typename TA, typename TV>
__global__ void __launch_bounds__(1024) gxr_kernel(const __grid_constant__ GridXfer X, const TA* __restrict__ tabg,
                                                   const TV* __restrict__ rf, TV* __restrict__ rc) {
    __shared__ TA tab[3 * 27];
    const int N0 = X.N[0], N1 = X.N[1], N2 = X.N[2];
    const long long S = X.n[0], S2 = (long long)X.n[0] * X.n[1];
    const int nlines = N1 * X.nk;                         // the rows of this matrix: coarse planes k0 .. k0 + nk - 1
    int cur = -1;
    for (int line = blockIdx.x; line < nlines; line += gridDim.x) {
        const int Kl = line / N1, J = line - Kl * N1, K = X.k0 + Kl;
        const int cy = gx_class(J, N1), cz = gx_class(K, N2);
        if (3 * cy + 9 * cz != cur) {          // CTA-uniform: the three x classes of this line's (y, z) class
            __syncthreads();
            cur = 3 * cy + 9 * cz;
            for (int i = threadIdx.x; i < 81; i += blockDim.x) tab[i] = tabg[(cur + i / 27) * 27 + i % 27];
            __syncthreads();
        }
        const int ay = gx_allowed(cy, N1), az = gx_allowed(cz, N2);
        const TV* f0 = rf + (S2 * (2LL * K) + S * (2LL * J) - X.shift);
        TV* out = rc + (long long)line * N0;
        const bool inner_line = ay == 7 && az == 7;          // CTA-uniform
        for (int I0 = 0; I0 < N0; I0 += blockDim.x) {
            const int I = I0 + threadIdx.x;
            const bool act = I < N0;
            const int cx = gx_class(act ? I : 1, N0);
            // warps of interior nodes of an interior line: all 27 products, no predicates (the mask arguments are
            // literal constants, so the tests fold away)
            const bool plain = __all_sync(0xffffffffu, !act || (inner_line && cx == 1 && N0 > 2));
            if (!act) continue;
            if (plain) out[I] = gxr_row<TA, TV>(tab + 27, f0 + 2 * I, S, S2, 7, 7, 7);
            else out[I] = gxr_row<TA, TV>(tab + cx * 27, f0 + 2 * I, S, S2, gx_allowed(cx, N0), ay, az);
        }
    }
}

// ---- block variants (nrhs > 1): blocks are stored RHS-fastest, a node is one contiguous segment of m values ---------------------
// One warp = one row at a time, lane j = right-hand side j (+ 32, ... when m > 32): every access is a coalesced segment
// (256 bytes for m = 32 Float64), the class tables sit in shared memory, no matrix stream.  The CSR block kernel these
// replace reads 12 bytes per non-zero and gathers through L2 (cfg4: restriction 445 us, prolongation 372 us,
// profiles/r02o_cfg4.json).  Products in stored order per (row, j): bit-identical to csr_stream_mrhs_kernel.
template <typename TA, typename TV>
__global__ void __launch_bounds__(256, 3) gxp_mrhs_kernel(const __grid_constant__ GridXfer X, int m, const TA* __restrict__ tabg,
                                                       const TV* __restrict__ xc, TV* __restrict__ xf) {
    __shared__ TA tab[GXP_TAB];
    for (int i = threadIdx.x; i < GXP_TAB; i += blockDim.x) tab[i] = tabg[i];
    __syncthreads();
    const int n0 = X.n[0], n1 = X.n[1], N0 = X.N[0], N1 = X.N[1];
    const long long cs2 = (long long)N0 * N1;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int nlines = n1 * X.nk;
    for (int line = blockIdx.x; line < nlines; line += gridDim.x) {
        const int kl = line / n1, jj = line - kl * n1, k = X.k0 + kl, b = jj & 1, c = k & 1;
        const long long q = ((long long)(k >> 1) * N1 + (jj >> 1)) * N0 - X.shift;        // coarse node (0, j/2, k/2)
        for (int i = w; i < n0; i += nw) {
            const int a = i & 1;
            const TA* tv = tab + (a + 2 * b + 4 * c) * 8;
            const long long row = (long long)line * n0 + i, p0 = q + (i >> 1);
            for (int j = lane; j < m; j += 32) {
                TV acc = VT<TV>::zero();
                int idx = 0;
#pragma unroll
                for (int dz = 0; dz < 2; ++dz) {
                    if (dz > c) break;
#pragma unroll
                    for (int dy = 0; dy < 2; ++dy) {
                        if (dy > b) break;
                        const long long p = p0 + dy * N0 + dz * cs2;
                        acc = acc + tv[idx] * ldg_(xc + (p * m + j));
                        ++idx;
                        if (a) {
                            acc = acc + tv[idx] * ldg_(xc + ((p + 1) * m + j));
                            ++idx;
                        }
                    }
                }
                xf[row * m + j] = xf[row * m + j] + acc;
            }
        }
    }
}
template <typename TA, typename TV>
__global__ void __launch_bounds__(256, 3) gxr_mrhs_kernel(const __grid_constant__ GridXfer X, int m, const TA* __restrict__ tabg,
                                                       const TV* __restrict__ rf, TV* __restrict__ rc) {
    __shared__ TA tab[GXR_TAB];
    for (int i = threadIdx.x; i < GXR_TAB; i += blockDim.x) tab[i] = tabg[i];
    __syncthreads();
    const int N0 = X.N[0], N1 = X.N[1], N2 = X.N[2];
    const long long S = X.n[0], S2 = (long long)X.n[0] * X.n[1];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int nlines = N1 * X.nk;
    for (int line = blockIdx.x; line < nlines; line += gridDim.x) {
        const int Kl = line / N1, J = line - Kl * N1, K = X.k0 + Kl;
        const int cy = gx_class(J, N1), cz = gx_class(K, N2);
        const int ay = gx_allowed(cy, N1), az = gx_allowed(cz, N2);
        const long long f0 = S2 * (2LL * K) + S * (2LL * J) - X.shift;
        for (int I = w; I < N0; I += nw) {
            const int cx = gx_class(I, N0), ax = gx_allowed(cx, N0);
            const TA* tv = tab + (cx + 3 * cy + 9 * cz) * 27;
            const long long f = f0 + 2 * I, row = (long long)line * N0 + I;
            const bool plain = ax == 7 && ay == 7 && az == 7;      // warp-uniform: interior node, all 27 products
            for (int j = lane; j < m; j += 32) {
                TV acc = VT<TV>::zero();
                if (plain) {
#pragma unroll
                    for (int e = 0; e < 27; ++e)
                        acc = acc + tv[e] * ldg_(rf + ((f + (e / 9 - 1) * S2 + ((e / 3) % 3 - 1) * S + (e % 3 - 1)) * m + j));
                } else {
#pragma unroll 1
                    for (int e = 0; e < 27; ++e) {
                        const int dz = e / 9 - 1, dy = (e / 3) % 3 - 1, dx = e % 3 - 1;
                        if (((az >> (dz + 1)) & 1) && ((ay >> (dy + 1)) & 1) && ((ax >> (dx + 1)) & 1))
                            acc = acc + tv[e] * ldg_(rf + ((f + dz * S2 + dy * S + dx) * m + j));
                    }
                }
                rc[row * m + j] = acc;
            }
        }
    }
}

}  // namespace mgb200
