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

// step k, part 1 (one CTA): pivot search in column k (first maximum of |a_ik|, as LAPACK),
// row swap over all columns, multipliers a_ik /= a_kk.
template <typename TV>
__global__ void __launch_bounds__(256) lu_pivot_kernel(TV* __restrict__ a, int n, int k, int* __restrict__ piv,
                                                       int* __restrict__ info) {
    __shared__ double smax[256];
    __shared__ int sidx[256];
    __shared__ int sp;
    const int tid = threadIdx.x;
    double best = -1.0;
    int bi = k;
    for (int i = k + tid; i < n; i += 256) {
        double v = abs2(a[(size_t)i * n + k]);
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
            double o = smax[tid + s];
            int oi = sidx[tid + s];
            if (o > smax[tid] || (o == smax[tid] && oi < sidx[tid])) {
                smax[tid] = o;
                sidx[tid] = oi;
            }
        }
        __syncthreads();
    }
    if (tid == 0) {
        sp = sidx[0];
        piv[k] = sp;
        if (!(smax[0] > 0.0)) atomicExch(info, k + 1);  // exactly singular pivot
    }
    __syncthreads();
    const int p = sp;
    if (p != k) {
        for (int j = tid; j < n; j += 256) {
            TV t = a[(size_t)k * n + j];
            a[(size_t)k * n + j] = a[(size_t)p * n + j];
            a[(size_t)p * n + j] = t;
        }
    }
    __syncthreads();
    const TV akk = a[(size_t)k * n + k];
    for (int i = k + 1 + tid; i < n; i += 256) a[(size_t)i * n + k] = a[(size_t)i * n + k] / akk;
}

__device__ __forceinline__ double operator_div(double a, double b) { return a / b; }

// step k, part 2: trailing update a_ij -= a_ik * a_kj, i,j > k
template <typename TV>
__global__ void lu_update_kernel(TV* __restrict__ a, int n, int k) {
    const int j = k + 1 + blockIdx.x * blockDim.x + threadIdx.x;
    const int i = k + 1 + blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= n || j >= n) return;
    a[(size_t)i * n + j] = a[(size_t)i * n + j] - a[(size_t)i * n + k] * a[(size_t)k * n + j];
}

// Linv (unit lower) and Uinv (upper), row-major n x n, zero outside their triangles.
// One thread per column j of the inverse; all threads walk rows in lock step so the factor
// entries are broadcast loads and the partial solutions are coalesced.
template <typename TV>
__global__ void lower_inverse_kernel(const TV* __restrict__ lu, int n, TV* __restrict__ linv) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    for (int i = 0; i < n; ++i) {
        TV y;
        if (i < j) {
            y = VT<TV>::zero();
        } else if (i == j) {
            y = VT<TV>::one();
        } else {
            TV s = VT<TV>::zero();
            for (int k = j; k < i; ++k) s = s + lu[(size_t)i * n + k] * linv[(size_t)k * n + j];
            y = -s;
        }
        linv[(size_t)i * n + j] = y;
    }
}

template <typename TV>
__global__ void upper_inverse_kernel(const TV* __restrict__ lu, int n, TV* __restrict__ uinv) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    for (int i = n - 1; i >= 0; --i) {
        TV y;
        if (i > j) {
            y = VT<TV>::zero();
        } else {
            TV s = (i == j) ? VT<TV>::one() : VT<TV>::zero();
            for (int k = i + 1; k <= j; ++k) s = s - lu[(size_t)i * n + k] * uinv[(size_t)k * n + j];
            y = s / lu[(size_t)i * n + i];
        }
        uinv[(size_t)i * n + j] = y;
    }
}

// y[i*m+c] = sum_{j<=i} Linv[i][j] * b[perm[j]*m+c]      (one warp per (row, rhs))
template <typename TW, typename TV>
__global__ void lower_apply_kernel(int n, int m, const TW* __restrict__ linv, const int* __restrict__ perm,
                                   const TV* __restrict__ b, TW* __restrict__ y) {
    const int lane = threadIdx.x & 31;
    const long long w = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    if (w >= (long long)n * m) return;
    const int i = (int)(w / m), c = (int)(w % m);
    TW acc = VT<TW>::zero();
    for (int j = lane; j <= i; j += 32) acc = acc + linv[(size_t)i * n + j] * widen(b[(size_t)perm[j] * m + c]);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc = acc + shfl_xor_(acc, s);
    if (lane == 0) y[(size_t)i * m + c] = acc;
}

// x[i*m+c] = sum_{j>=i} Uinv[i][j] * y[j*m+c]
template <typename TW, typename TV>
__global__ void upper_apply_kernel(int n, int m, const TW* __restrict__ uinv, const TW* __restrict__ y,
                                   TV* __restrict__ x) {
    const int lane = threadIdx.x & 31;
    const long long w = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    if (w >= (long long)n * m) return;
    const int i = (int)(w / m), c = (int)(w % m);
    TW acc = VT<TW>::zero();
    for (int j = i + lane; j < n; j += 32) acc = acc + uinv[(size_t)i * n + j] * y[(size_t)j * m + c];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc = acc + shfl_xor_(acc, s);
    if (lane == 0) narrow(acc, x[(size_t)i * m + c]);
}

}  // namespace mgb200
