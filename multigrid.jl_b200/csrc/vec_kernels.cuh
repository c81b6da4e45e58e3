// Vector (BLAS-1 / tall-skinny BLAS-2) kernels of the cycle and the Krylov drivers.
//
// Reference call sites replaced:
//   norm / dot                      SolveFuncs.jl:15-30, FGMRES.jl:64,97; KrylovMethods cg/fgmres
//   BLAS.axpy! (addVectors)         SpMatMul.jl:29-36
//   BLAS.gemv!('C'/'N') on the basis FGMRES.jl:95,121; KrylovMethods.fgmres Gram-Schmidt
//   x .+= d.*r with x = 0           MGcycle.jl:129 (first sweep of a cycle)
//
// All reductions are deterministic: per-CTA partial sums are combined in a fixed order by
// the last CTA to finish (atomic ticket), so repeated runs give identical bits.
#pragma once
#include "common.cuh"

namespace mgb200 {

constexpr int RED_THREADS = 256;
constexpr int RED_MAX_BLOCKS = 1184;  // 148 SMs x 8
constexpr int MAXK = 32;              // max Krylov basis columns handled by the fused kernels

struct ReduceWs {
    double* partials;   // RED_MAX_BLOCKS * 2*(MAXK+2) doubles
    unsigned* counter;  // ticket
};

// Block-reduce NV doubles per thread, publish the per-CTA partials, and let the last CTA
// combine them in block order into out[0..NV).
template <int NV>
__device__ __forceinline__ void grid_reduce_store(double (&v)[NV], ReduceWs ws, double* __restrict__ out) {
    __shared__ double sred[RED_THREADS / 32][NV];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double a = v[i];
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) a += __shfl_down_sync(0xffffffffu, a, s);
        if (lane == 0) sred[warp][i] = a;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double a = 0.0;
#pragma unroll
        for (int w = 0; w < RED_THREADS / 32; ++w) a += sred[w][threadIdx.x];
        ws.partials[(size_t)blockIdx.x * NV + threadIdx.x] = a;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = atomicAdd(ws.counter, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        if (threadIdx.x < NV) {
            double a = 0.0;
            for (unsigned bIdx = 0; bIdx < gridDim.x; ++bIdx)
                a += __ldcg(ws.partials + (size_t)bIdx * NV + threadIdx.x);
            out[threadIdx.x] = a;
        }
        if (threadIdx.x == 0) *ws.counter = 0u;
    }
}

// out[0..1] = sum conj(x_i) * y_i  (re, im)
template <typename TV>
__global__ void __launch_bounds__(RED_THREADS) dot_kernel(long long n, const TV* __restrict__ x,
                                                          const TV* __restrict__ y, ReduceWs ws,
                                                          double* __restrict__ out) {
    double v[2] = {0.0, 0.0};
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        TV p = conj_(x[i]) * y[i];
        v[0] += VT<TV>::re(p);
        v[1] += VT<TV>::im(p);
    }
    grid_reduce_store<2>(v, ws, out);
}

// out[0] = sum |x_i|^2
template <typename TV>
__global__ void __launch_bounds__(RED_THREADS) norm2sq_kernel(long long n, const TV* __restrict__ x,
                                                              ReduceWs ws, double* __restrict__ out) {
    double v[1] = {0.0};
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        v[0] += abs2(x[i]);
    grid_reduce_store<1>(v, ws, out);
}

// out[2c], out[2c+1] = sum conj(V[i + c*ld]) * w[i]  for c < kk <= K   (t = V^H w, one pass over w)
template <typename TV, int K>
__global__ void __launch_bounds__(RED_THREADS) multi_dot_kernel(long long n, const TV* __restrict__ V,
                                                                long long ld, int kk,
                                                                const TV* __restrict__ w, ReduceWs ws,
                                                                double* __restrict__ out) {
    double v[2 * K];
#pragma unroll
    for (int c = 0; c < 2 * K; ++c) v[c] = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const TV wi = w[i];
#pragma unroll
        for (int c = 0; c < K; ++c) {
            if (c < kk) {
                TV p = conj_(V[i + c * ld]) * wi;
                v[2 * c] += VT<TV>::re(p);
                v[2 * c + 1] += VT<TV>::im(p);
            }
        }
    }
    grid_reduce_store<2 * K>(v, ws, out);
}

template <typename TV>
struct Coefs {
    TV c[MAXK];
};

// w[i] = beta*w[i] + sum_c coef[c] * V[i + c*ld]; optionally out[0] = sum |w_i|^2 of the result
template <typename TV, bool WITH_NORM>
__global__ void __launch_bounds__(RED_THREADS) multi_axpy_kernel(long long n, const TV* __restrict__ V,
                                                                 long long ld, int K, Coefs<TV> coef,
                                                                 double beta, TV* __restrict__ w,
                                                                 ReduceWs ws, double* __restrict__ out) {
    double v[1] = {0.0};
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        TV a = (beta == 0.0) ? VT<TV>::zero() : beta * w[i];
        for (int c = 0; c < K; ++c) a = a + coef.c[c] * V[i + c * ld];
        w[i] = a;
        if (WITH_NORM) v[0] += abs2(a);
    }
    if (WITH_NORM) grid_reduce_store<1>(v, ws, out);
}

// y = a*x + b*y  (host scalars);  b == 0 does not read y
template <typename TV>
__global__ void axpby_kernel(long long n, TV a, const TV* x, TV b, TV* y, int b_is_zero) {  // x may alias y
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        y[i] = b_is_zero ? a * x[i] : (b * y[i] + a * x[i]);
}

// BiCGStab updates (KrylovMethods.bicgstb):  p = r + beta*(p - omega*v)   and   x += alpha*phat + omega*shat
template <typename TV>
__global__ void bicg_p_kernel(long long n, TV beta, TV omega, const TV* __restrict__ r, const TV* __restrict__ v,
                              TV* __restrict__ p) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        p[i] = r[i] + beta * (p[i] - omega * v[i]);
}
template <typename TV>
__global__ void bicg_x_kernel(long long n, TV alpha, const TV* __restrict__ phat, TV omega,
                              const TV* __restrict__ shat, TV* __restrict__ x, int with_shat) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        x[i] = with_shat ? x[i] + (alpha * phat[i] + omega * shat[i]) : x[i] + alpha * phat[i];
}

// first sweep of a cycle from x = 0:  x = 0 + d .* b   (MGcycle.jl:28-31 skipped, :129)
template <typename TV>
__global__ void diag_scale_kernel(long long n, int m, const TV* __restrict__ d, const TV* __restrict__ b,
                                  TV* __restrict__ x) {
    const long long total = n * m;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x)
        x[i] = VT<TV>::zero() + d[i / m] * b[i];
}

// column-major n x m (the reference's host layout)  <->  RHS-fastest (device layout)
template <typename TV>
__global__ void colmajor_to_rhsfast_kernel(long long n, int m, const TV* __restrict__ in, TV* __restrict__ out) {
    const long long total = n * m;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        long long i = t / m;
        int j = (int)(t % m);
        out[t] = in[i + (long long)j * n];
    }
}
template <typename TV>
__global__ void rhsfast_to_colmajor_kernel(long long n, int m, const TV* __restrict__ in, TV* __restrict__ out) {
    const long long total = n * m;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        long long i = t % n;
        int j = (int)(t / n);
        out[t] = in[i * m + j];
    }
}

// ---------------------------------------------------------------------------------------------
// block-Krylov pieces (KrylovMethods.blockCG, SolveFuncs.jl:113): n x m blocks stored RHS-fastest.
// ---------------------------------------------------------------------------------------------
constexpr int GRAM_ROWS = 32;      // rows staged per tile
constexpr int GRAM_THREADS = 256;
constexpr int GRAM_MAX_BLOCKS = 592;

// G = X^H Y  (m x m, row-major, interleaved re/im):  G[a][b] = sum_i conj(X[i,a]) * Y[i,b].
// Tiles of GRAM_ROWS rows of X and Y are staged in shared memory; each thread owns the (a,b) pairs
// p = tid, tid+256, ...  Per-CTA partial Gram matrices are combined in block order by the last CTA.
template <typename TV>
__global__ void __launch_bounds__(GRAM_THREADS) gram_kernel(long long n, int m, const TV* __restrict__ X,
                                                            const TV* __restrict__ Y, double* __restrict__ partials,
                                                            unsigned* __restrict__ counter, double* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char gsm[];
    TV* xs = reinterpret_cast<TV*>(gsm);
    TV* ys = xs + (size_t)GRAM_ROWS * m;
    __shared__ bool is_last;
    const int mm = m * m;
    constexpr int MAXP = 16;  // pairs per thread: m <= 64
    double accr[MAXP], acci[MAXP];
#pragma unroll
    for (int q = 0; q < MAXP; ++q) accr[q] = acci[q] = 0.0;
    const long long ntiles = (n + GRAM_ROWS - 1) / GRAM_ROWS;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const long long r0 = t * GRAM_ROWS;
        const int rows = (int)min((long long)GRAM_ROWS, n - r0);
        __syncthreads();
        for (int e = threadIdx.x; e < rows * m; e += GRAM_THREADS) {
            xs[e] = X[r0 * m + e];
            ys[e] = Y[r0 * m + e];
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < MAXP; ++q) {
            const int p = threadIdx.x + q * GRAM_THREADS;
            if (p < mm) {
                const int a = p / m, b = p % m;
                double sr = accr[q], si = acci[q];
                for (int r = 0; r < rows; ++r) {
                    TV pr = conj_(xs[r * m + a]) * ys[r * m + b];
                    sr += VT<TV>::re(pr);
                    si += VT<TV>::im(pr);
                }
                accr[q] = sr;
                acci[q] = si;
            }
        }
    }
#pragma unroll
    for (int q = 0; q < MAXP; ++q) {
        const int p = threadIdx.x + q * GRAM_THREADS;
        if (p < mm) {
            partials[((size_t)blockIdx.x * mm + p) * 2] = accr[q];
            partials[((size_t)blockIdx.x * mm + p) * 2 + 1] = acci[q];
        }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned tk = atomicAdd(counter, 1u);
        is_last = (tk == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        for (int p = threadIdx.x; p < 2 * mm; p += GRAM_THREADS) {
            double a = 0.0;
            for (unsigned bIdx = 0; bIdx < gridDim.x; ++bIdx) a += __ldcg(partials + (size_t)bIdx * mm * 2 + p);
            out[p] = a;
        }
        if (threadIdx.x == 0) *counter = 0u;
    }
}

// out[i,b] = base[i,b] + sum_a X[i,a] * C[a,b]   (C: m x m row-major on the device; out may alias base,
// never X).  base == nullptr means 0.
template <typename TV>
__global__ void block_axpy_kernel(long long n, int m, const TV* __restrict__ X, const TV* __restrict__ C,
                                  const TV* base, TV* out) {
    extern __shared__ __align__(16) unsigned char csm[];
    TV* cs = reinterpret_cast<TV*>(csm);
    for (int e = threadIdx.x; e < m * m; e += blockDim.x) cs[e] = C[e];
    __syncthreads();
    const long long total = n * m;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const long long i = t / m;
        const int b = (int)(t % m);
        TV acc = base ? base[t] : VT<TV>::zero();
        const TV* xr = X + i * m;
        for (int a = 0; a < m; ++a) acc = acc + xr[a] * cs[a * m + b];
        out[t] = acc;
    }
}

// out[b] = sum_i |R[i,b]|^2 for b < m  (deterministic two-stage reduction, m <= 64)
template <typename TV>
__global__ void __launch_bounds__(RED_THREADS) colnorm2_kernel(long long n, int m, const TV* __restrict__ R,
                                                               double* __restrict__ partials,
                                                               unsigned* __restrict__ counter, double* __restrict__ out) {
    __shared__ double sacc[RED_THREADS];
    __shared__ bool is_last;
    // thread -> column b = tid % mp, row phase = tid / mp
    int mp = 1;
    while (mp < m) mp <<= 1;
    const int rows_per_pass = RED_THREADS / mp;
    const int b = threadIdx.x % mp, ph = threadIdx.x / mp;
    double a = 0.0;
    if (b < m && rows_per_pass > 0) {
        for (long long i = (long long)blockIdx.x * rows_per_pass + ph; i < n; i += (long long)gridDim.x * rows_per_pass)
            a += abs2(R[i * m + b]);
    }
    sacc[threadIdx.x] = a;
    __syncthreads();
    if (threadIdx.x < mp) {
        double s = 0.0;
        for (int q = 0; q < rows_per_pass; ++q) s += sacc[q * mp + threadIdx.x];
        if (threadIdx.x < m) partials[(size_t)blockIdx.x * m + threadIdx.x] = s;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned tk = atomicAdd(counter, 1u);
        is_last = (tk == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        if (threadIdx.x < m) {
            double s = 0.0;
            for (unsigned bIdx = 0; bIdx < gridDim.x; ++bIdx) s += __ldcg(partials + (size_t)bIdx * m + threadIdx.x);
            out[threadIdx.x] = s;
        }
        if (threadIdx.x == 0) *counter = 0u;
    }
}

// ---------------------------------------------------------------------------------------------
// PCG pieces with device-resident scalars (no host round trip between the kernels)
//   scal[0..1] = gamma = <r,z>, scal[2..3] = delta = <p,Ap>
// cg_update:  alpha = gamma/delta; if alpha is Inf or negative nothing is changed (the host
//             sees the same scalars afterwards and stops with flag -2, as KrylovMethods.cg does);
//             else x += alpha p, r -= alpha Ap and out[0] = ||r||^2.
// ---------------------------------------------------------------------------------------------
template <typename TV>
__global__ void __launch_bounds__(RED_THREADS) cg_update_kernel(long long n, const double* __restrict__ scal,
                                                                const TV* __restrict__ p,
                                                                const TV* __restrict__ Ap, TV* __restrict__ x,
                                                                TV* __restrict__ r, ReduceWs ws,
                                                                double* __restrict__ out) {
    const double gamma = scal[0], delta = scal[2];
    const double alpha = gamma / delta;
    const bool bad = isinf(alpha) || (alpha < 0.0);
    double v[1] = {0.0};
    if (!bad) {
        for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
             i += (long long)gridDim.x * blockDim.x) {
            x[i] = x[i] + alpha * p[i];
            TV ri = r[i] + (-alpha) * Ap[i];
            r[i] = ri;
            v[0] += abs2(ri);
        }
    }
    grid_reduce_store<1>(v, ws, out);
}

// p = z + beta*p with beta = <z,r>_new / gamma_old  (scal! followed by axpy! in the package)
template <typename TV>
__global__ void cg_direction_kernel(long long n, const double* __restrict__ gamma_new,
                                    const double* __restrict__ gamma_old, const TV* __restrict__ z,
                                    TV* __restrict__ p) {
    const double beta = gamma_new[0] / gamma_old[0];
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        p[i] = z[i] + beta * p[i];
}

// precision change of a vector block (mixed-precision preconditioner, SolveFuncs.jl:57: bl[:] .= b ... z2[:] .= z)
__device__ __forceinline__ void convert_val(double a, float& o) { o = (float)a; }
__device__ __forceinline__ void convert_val(float a, double& o) { o = (double)a; }
__device__ __forceinline__ void convert_val(cplx a, cplxf& o) { o = make_cplxf((float)a.x, (float)a.y); }
__device__ __forceinline__ void convert_val(cplxf a, cplx& o) { o = make_cplx((double)a.x, (double)a.y); }
template <typename TS, typename TD>
__global__ void convert_kernel(long long n, const TS* __restrict__ src, TD* __restrict__ dst) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) convert_val(src[i], dst[i]);
}

}  // namespace mgb200
