// Coarsest-grid solver on the device: dense LU with partial pivoting, factorised once.
//
// Replaces `param.LU = lu(sparse(AT'))` (src/Multigrid/MGsetup.jl:350, UMFPACK) and
// `z = param.LU\b` (src/Multigrid/MGcycle.jl:176-179).  Two dependent triangular solves are
// a chain of ~2n serial steps, which is latency-bound on a GPU and sits on the critical path
// of every cycle.  So after the factorisation P A = L U the two triangular factors are
// inverted once (column by column, each column an independent substitution), and a coarsest
// solve becomes two dense triangular matrix-vector products x = U^-1 (L^-1 (P b)) with no
// serial dependency.  No CPU fallback: factorisation, inversion and solves all run on device.
#pragma once
#include "common.cuh"

namespace mgb200 {

template <typename TV, typename TW>
__global__ void densify_kernel(int n, const int* __restrict__ rowptr, const int* __restrict__ colind,
                               const TV* __restrict__ val, TW* __restrict__ a) {
    int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    for (int k = rowptr[row]; k < rowptr[row + 1]; ++k) {
        // duplicate entries cannot occur in a CSC matrix coming from the reference; plain store
        a[(size_t)row * n + colind[k]] = widen(val[k]);
    }
}

// ---- blocked right-looking LU with partial pivoting (panels of LU_NB columns) ------------------------------------------------
// Per panel: (1) lu_panel_kernel - a cooperative launch factorises the panel column by column (pivot search over all
// CTAs, two grid barriers per column); (2) lu_swap_kernel applies the panel's row interchanges to the columns outside
// it; (3) lu_trsm_kernel solves L11 U12 = A12; (4) lu_gemm_kernel updates the trailing matrix A22 -= L21 U12 from
// shared-memory tiles.  4 launches per 32 columns instead of 2 per column, and the trailing matrix is streamed once per
// panel instead of once per column (the unblocked rank-1 form of round 1 needed 110 s for n = 23 376).
constexpr int LU_NB = 32;

// c - a * b (the trailing update is the one place of the library where a fused multiply-add is used: the factorisation
// has no bit-level counterpart in the reference, whose UMFPACK / LAPACK kernels contract as they please)
__device__ __forceinline__ double fnma_(double a, double b, double c) { return __fma_rn(-a, b, c); }
__device__ __forceinline__ cplx fnma_(cplx a, cplx b, cplx c) {
    return make_cplx(__fma_rn(a.y, b.y, __fma_rn(-a.x, b.x, c.x)), __fma_rn(-a.y, b.x, __fma_rn(-a.x, b.y, c.y)));
}

struct LuPivot {
    double val;
    int idx;
};
// loads that bypass the (non-coherent) L1: the panel kernel reads what other CTAs wrote before the last grid barrier
__device__ __forceinline__ double ld_cg_(const double* p) { return __ldcg(p); }
__device__ __forceinline__ cplx ld_cg_(const cplx* p) {
    const double2 v = __ldcg(reinterpret_cast<const double2*>(p));
    return make_cplx(v.x, v.y);
}
// grid barrier of a cooperative launch: arrive on a monotonically increasing counter (no reset, so no ABA problem)
__device__ __forceinline__ void lu_grid_sync(unsigned* counter, unsigned& epoch) {
    __syncthreads();
    if (threadIdx.x == 0) {
        epoch += gridDim.x;
        __threadfence();
        atomicAdd(counter, 1u);
        while (*reinterpret_cast<volatile unsigned*>(counter) < epoch) {}
        __threadfence();
    }
    __syncthreads();
}
// panel columns [k0, k0 + nb): all CTAs are resident (cooperative launch); rows are dealt to warps round-robin
template <typename TW>
__global__ void __launch_bounds__(256) lu_panel_kernel(TW* __restrict__ a, int n, int k0, int nb, int* __restrict__ piv,
                                                       int* __restrict__ info, LuPivot* __restrict__ cand, unsigned* counter,
                                                       unsigned epoch0) {
    __shared__ double smax[256];
    __shared__ int sidx[256];
    const int tid = threadIdx.x, lane = tid & 31;
    const int gw = (blockIdx.x * 256 + tid) >> 5, nw = gridDim.x * 8;      // global warp id, warps in the grid
    unsigned epoch = epoch0;
    for (int j = 0; j < nb; ++j) {
        const int col = k0 + j;
        // (i) pivot candidates: first maximum of |a_i,col| over this CTA's rows (thread-strided), then over the CTAs
        double best = -1.0;
        int bi = col;
        for (int i = col + blockIdx.x * 256 + tid; i < n; i += gridDim.x * 256) {
            const double v = abs2(ld_cg_(a + (size_t)i * n + col));
            if (v > best) {
                best = v;
                bi = i;
            }
        }
        smax[tid] = best;
        sidx[tid] = bi;
        __syncthreads();
        for (int s = 128; s > 0; s >>= 1) {
            if (tid < s) {
                const double o = smax[tid + s];
                const int oi = sidx[tid + s];
                if (o > smax[tid] || (o == smax[tid] && oi < sidx[tid])) {
                    smax[tid] = o;
                    sidx[tid] = oi;
                }
            }
            __syncthreads();
        }
        if (tid == 0) {
            cand[blockIdx.x].val = smax[0];
            cand[blockIdx.x].idx = sidx[0];
        }
        lu_grid_sync(counter, epoch);
        // every CTA reduces the candidates the same way (ties: smallest row, as LAPACK's idamax)
        double gb = -1.0;
        int p = col;
        for (int c = 0; c < (int)gridDim.x; ++c) {
            const double v = __ldcg(&cand[c].val);
            const int vi = __ldcg(&cand[c].idx);
            if (v > gb || (v == gb && vi < p)) {
                gb = v;
                p = vi;
            }
        }
        if (blockIdx.x == 0) {
            if (tid == 0) {
                piv[col] = p;
                if (!(gb > 0.0)) atomicExch(info, col + 1);      // exactly singular pivot
            }
            if (p != col && tid < nb) {                           // interchange inside the panel
                const TW t = ld_cg_(a + (size_t)col * n + k0 + tid);
                a[(size_t)col * n + k0 + tid] = ld_cg_(a + (size_t)p * n + k0 + tid);
                a[(size_t)p * n + k0 + tid] = t;
            }
        }
        lu_grid_sync(counter, epoch);
        // (iii) multipliers and the rank-1 update of the panel's remaining columns: one warp per row, lanes = columns
        const TW akk = ld_cg_(a + (size_t)col * n + col);
        const TW ukc = (lane > j && lane < nb) ? ld_cg_(a + (size_t)col * n + k0 + lane) : VT<TW>::zero();
        for (int i = col + 1 + gw; i < n; i += nw) {
            TW* row = a + (size_t)i * n + k0;
            const TW l = ld_cg_(row + j) / akk;
            const TW v = (lane < nb) ? ld_cg_(row + lane) : VT<TW>::zero();
            __syncwarp();
            if (lane == j) row[lane] = l;
            else if (lane > j && lane < nb) row[lane] = v - l * ukc;
        }
        // the next column's pivot search reads what other CTAs just updated
        lu_grid_sync(counter, epoch);
    }
}
// the panel's interchanges applied to every column outside the panel
template <typename TW>
__global__ void lu_swap_kernel(TW* __restrict__ a, int n, int k0, int nb, const int* __restrict__ piv) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n - nb) return;
    if (j >= k0) j += nb;
    for (int s = 0; s < nb; ++s) {
        const int r = k0 + s, p = piv[r];
        if (p != r) {
            const TW t = a[(size_t)r * n + j];
            a[(size_t)r * n + j] = a[(size_t)p * n + j];
            a[(size_t)p * n + j] = t;
        }
    }
}
// U12 = L11^-1 A12: one thread per column right of the panel, forward substitution with the unit lower block in shared memory
template <typename TW>
__global__ void __launch_bounds__(128) lu_trsm_kernel(TW* __restrict__ a, int n, int k0, int nb) {
    __shared__ TW L11[LU_NB * LU_NB];
    for (int i = threadIdx.x; i < nb * nb; i += blockDim.x) L11[i] = a[(size_t)(k0 + i / nb) * n + k0 + i % nb];
    __syncthreads();
    const int j = k0 + nb + blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    TW xcol[LU_NB];
#pragma unroll
    for (int r = 0; r < LU_NB; ++r) {
        if (r < nb) {
            TW v = a[(size_t)(k0 + r) * n + j];
#pragma unroll
            for (int s2 = 0; s2 < LU_NB; ++s2)
                if (s2 < r) v = v - L11[r * nb + s2] * xcol[s2];
            xcol[r] = v;
            a[(size_t)(k0 + r) * n + j] = v;
        }
    }
}
// A22 -= L21 U12: T x T tiles (64 for 8-byte values, 32 for complex), 256 threads, (T/16)^2 outputs per thread, the
// nb-deep operands in shared memory
template <typename TW>
struct LuTile {
    static constexpr int T = sizeof(TW) == 8 ? 64 : 32;
};
template <typename TW>
__global__ void __launch_bounds__(256) lu_gemm_kernel(TW* __restrict__ a, int n, int k0, int nb) {
    constexpr int T = LuTile<TW>::T, R = T / 16;
    __shared__ TW As[T][LU_NB + 1];
    __shared__ TW Bs[LU_NB][T];
    const int r0 = k0 + nb + blockIdx.y * T, c0 = k0 + nb + blockIdx.x * T;
    for (int i = threadIdx.x; i < T * nb; i += 256) {
        const int r = i / nb, c = i % nb;
        As[r][c] = (r0 + r < n) ? a[(size_t)(r0 + r) * n + k0 + c] : VT<TW>::zero();
    }
    for (int i = threadIdx.x; i < nb * T; i += 256) {
        const int r = i / T, c = i % T;
        Bs[r][c] = (c0 + c < n) ? a[(size_t)(k0 + r) * n + c0 + c] : VT<TW>::zero();
    }
    __syncthreads();
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;       // outputs (ty + 16 i, tx + 16 j)
    TW acc[R][R];
#pragma unroll
    for (int i = 0; i < R; ++i)
#pragma unroll
        for (int j = 0; j < R; ++j) {
            const int r = r0 + ty + 16 * i, c = c0 + tx + 16 * j;
            acc[i][j] = (r < n && c < n) ? a[(size_t)r * n + c] : VT<TW>::zero();
        }
    for (int k = 0; k < nb; ++k) {
        TW av[R], bv[R];
#pragma unroll
        for (int i = 0; i < R; ++i) av[i] = As[ty + 16 * i][k];
#pragma unroll
        for (int j = 0; j < R; ++j) bv[j] = Bs[k][tx + 16 * j];
#pragma unroll
        for (int i = 0; i < R; ++i)
#pragma unroll
            for (int j = 0; j < R; ++j) acc[i][j] = fnma_(av[i], bv[j], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < R; ++i)
#pragma unroll
        for (int j = 0; j < R; ++j) {
            const int r = r0 + ty + 16 * i, c = c0 + tx + 16 * j;
            if (r < n && c < n) a[(size_t)r * n + c] = acc[i][j];
        }
}

// ---- blocked inversion of the triangular factors ------------------------------------------------------------------------------
// X = T^-1 for T = L (unit lower, UPPER = false) or U (upper): block rows of TI_NB rows in dependency order
// (top-down for L, bottom-up for U).  The diagonal block is inverted by one CTA (tri_diag_inv_kernel); the rest of the
// block row is  X[R, C] = -X[R,R] (T[R, K] X[K, C])  over the block rows K already done (tri_inv_row_kernel: tiles of
// TI_NB x 64, operands from shared memory).
constexpr int TI_NB = 32;
template <typename TW, bool UPPER>
__global__ void __launch_bounds__(TI_NB) tri_diag_inv_kernel(const TW* __restrict__ lu, int n, int r0, int nbk, TW* __restrict__ x) {
    // thread j solves column j of the nbk x nbk block by substitution
    __shared__ TW Tb[TI_NB][TI_NB + 1];
    const int j = threadIdx.x;
    for (int i = 0; i < nbk; ++i)
        if (j < nbk) Tb[i][j] = lu[(size_t)(r0 + i) * n + r0 + j];
    __syncthreads();
    if (j >= nbk) return;
    TW col[TI_NB];
    if (!UPPER) {
        for (int i = 0; i < nbk; ++i) {
            TW s = (i == j) ? VT<TW>::one() : VT<TW>::zero();
            for (int k = j; k < i; ++k) s = s - Tb[i][k] * col[k];
            col[i] = (i < j) ? VT<TW>::zero() : s;           // unit diagonal
        }
    } else {
        for (int i = nbk - 1; i >= 0; --i) {
            TW s = (i == j) ? VT<TW>::one() : VT<TW>::zero();
            for (int k = i + 1; k <= j; ++k) s = s - Tb[i][k] * col[k];
            col[i] = (i > j) ? VT<TW>::zero() : s / Tb[i][i];
        }
    }
    for (int i = 0; i < nbk; ++i) x[(size_t)(r0 + i) * n + r0 + j] = col[i];
}
template <typename TW, bool UPPER>
__global__ void __launch_bounds__(256) tri_inv_row_kernel(const TW* __restrict__ lu, int n, int r0, int nbk, TW* __restrict__ x) {
    constexpr int TC = LuTile<TW>::T, CPT = TC / 8;
    __shared__ TW Ts[TI_NB][TI_NB + 1];      // T[R, K-chunk]; at the end X[R,R]
    __shared__ TW Xs[TI_NB][TC];             // X[K-chunk, C-tile]; at the end the product T[R,K] X[K,C]
    // column tile C: lower: columns [0, r0); upper: columns [r0 + nbk, n)
    const int c0 = UPPER ? r0 + nbk + blockIdx.x * TC : blockIdx.x * TC;
    const int cend = UPPER ? n : r0;
    const int tid = threadIdx.x;
    const int tr = tid >> 3, tc = (tid & 7) * CPT;      // thread: row tr (0..31), columns tc .. tc + CPT - 1
    TW acc[CPT];
#pragma unroll
    for (int q = 0; q < CPT; ++q) acc[q] = VT<TW>::zero();
    // K runs over the finished block rows that can be non-zero for this column tile:
    // lower: rows k in [c0, r0) (X[k, c] = 0 for k < c); upper: rows k in [r0 + nbk, min(n, c0 + TC)) (X[k, c] = 0 for k > c)
    const int kbeg = UPPER ? r0 + nbk : (c0 / TI_NB) * TI_NB;
    const int kend = UPPER ? (c0 + TC < n ? c0 + TC : n) : r0;
    for (int k0 = kbeg; k0 < kend; k0 += TI_NB) {
        __syncthreads();
        for (int i = tid; i < TI_NB * TI_NB; i += 256) {
            const int r = i / TI_NB, c = i % TI_NB;
            Ts[r][c] = (r < nbk && k0 + c < kend) ? lu[(size_t)(r0 + r) * n + k0 + c] : VT<TW>::zero();
        }
        for (int i = tid; i < TI_NB * TC; i += 256) {
            const int r = i / TC, c = i % TC;
            Xs[r][c] = (k0 + r < kend && c0 + c < cend) ? x[(size_t)(k0 + r) * n + c0 + c] : VT<TW>::zero();
        }
        __syncthreads();
#pragma unroll 4
        for (int k = 0; k < TI_NB; ++k) {
            const TW t = Ts[tr][k];
#pragma unroll
            for (int q = 0; q < CPT; ++q) acc[q] = acc[q] + t * Xs[k][tc + q];
        }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < CPT; ++q) Xs[tr][tc + q] = acc[q];
    for (int i = tid; i < TI_NB * TI_NB; i += 256) {
        const int r = i / TI_NB, c = i % TI_NB;
        Ts[r][c] = (r < nbk && c < nbk) ? x[(size_t)(r0 + r) * n + r0 + c] : VT<TW>::zero();
    }
    __syncthreads();
    // X[R, C] = -X[R,R] P
    TW out[CPT];
#pragma unroll
    for (int q = 0; q < CPT; ++q) out[q] = VT<TW>::zero();
    for (int k = 0; k < TI_NB; ++k) {
        const TW t = Ts[tr][k];
#pragma unroll
        for (int q = 0; q < CPT; ++q) out[q] = out[q] - t * Xs[k][tc + q];
    }
    if (tr < nbk)
#pragma unroll
        for (int q = 0; q < CPT; ++q)
            if (c0 + tc + q < cend) x[(size_t)(r0 + tr) * n + c0 + tc + q] = out[q];
}

// y[i*m+c] = sum_{j<=i} Linv[i][j] * b[perm[j]*m+c]      (one warp per row and chunk of up to AP_MC right-hand sides:
// the factor row is read once per chunk, not once per right-hand side)
constexpr int AP_MC = 8;
template <typename TW, typename TV, int MC>
__global__ void lower_apply_kernel(int n, int m, const TW* __restrict__ linv, const int* __restrict__ perm,
                                   const TV* __restrict__ b, TW* __restrict__ y) {
    const int lane = threadIdx.x & 31;
    const int nch = (m + MC - 1) / MC;
    const long long w = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    if (w >= (long long)n * nch) return;
    const int i = (int)(w / nch), c0 = (int)(w % nch) * MC;
    const int mc = m - c0 < MC ? m - c0 : MC;
    TW acc[MC];
#pragma unroll
    for (int c = 0; c < MC; ++c) acc[c] = VT<TW>::zero();
    for (int j = lane; j <= i; j += 32) {
        const TW l = linv[(size_t)i * n + j];
        const TV* bj = b + (size_t)perm[j] * m + c0;
#pragma unroll
        for (int c = 0; c < MC; ++c)
            if (c < mc) acc[c] = acc[c] + l * widen(bj[c]);
    }
#pragma unroll
    for (int c = 0; c < MC; ++c) {
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) acc[c] = acc[c] + shfl_xor_(acc[c], s);
        if (lane == 0 && c < mc) y[(size_t)i * m + c0 + c] = acc[c];
    }
}

// x[i*m+c] = sum_{j>=i} Uinv[i][j] * y[j*m+c]
template <typename TW, typename TV, int MC>
__global__ void upper_apply_kernel(int n, int m, const TW* __restrict__ uinv, const TW* __restrict__ y,
                                   TV* __restrict__ x) {
    const int lane = threadIdx.x & 31;
    const int nch = (m + MC - 1) / MC;
    const long long w = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    if (w >= (long long)n * nch) return;
    const int i = (int)(w / nch), c0 = (int)(w % nch) * MC;
    const int mc = m - c0 < MC ? m - c0 : MC;
    TW acc[MC];
#pragma unroll
    for (int c = 0; c < MC; ++c) acc[c] = VT<TW>::zero();
    for (int j = i + lane; j < n; j += 32) {
        const TW u = uinv[(size_t)i * n + j];
        const TW* yj = y + (size_t)j * m + c0;
#pragma unroll
        for (int c = 0; c < MC; ++c)
            if (c < mc) acc[c] = acc[c] + u * yj[c];
    }
#pragma unroll
    for (int c = 0; c < MC; ++c) {
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) acc[c] = acc[c] + shfl_xor_(acc[c], s);
        if (lane == 0 && c < mc) narrow(acc[c], x[(size_t)i * m + c0 + c]);
    }
}

}  // namespace mgb200
