// CSR row-gather kernels: y = op(A x) for the operators of the cycle.
//
// Reference call sites replaced:
//   SpMatMul(alpha,AT,x,beta,y)  src/Multigrid/SpMatMul.jl:4-26  with (alpha,beta) in
//   {(-1,1),(-1,0),(1,1),(1,0)}, the axpy that follows it (addVectors, :29-36) and the diagonal
//   relaxation update x .+= d.*r (MGcycle.jl:129,134) fused into one pass over the matrix.
//
// Design ("CSR-stream"): a CTA owns a contiguous block of rows.  The block's slice of the
// value and column-index arrays is contiguous in memory, so one elected thread moves it into
// shared memory with two TMA bulk copies (cp.async.bulk + mbarrier, no register staging,
// perfectly coalesced) while the other threads already fetch their row pointers and the
// b/d/x values of the epilogue.  Rows are then reduced from shared memory by 1..32 threads
// per row (chosen from the row-length statistics at upload); x is gathered through the
// read-only path and hits L1/L2 for stencil-like matrices.  With one thread per row the
// products are accumulated in stored order, exactly like the reference's row loop.
#pragma once
#include "common.cuh"

namespace mgb200 {

enum CsrMode { MODE_SPMV = 0, MODE_ADD = 1, MODE_RESID = 2, MODE_SWEEP = 3 };

// y-update shared by all kernels.  i = row*m + j (right-hand sides fastest).
template <int MODE, typename TV>
__device__ __forceinline__ void csr_epilogue(TV t, long long i, TV dval, const TV* __restrict__ x,
                                             const TV* __restrict__ b, TV* __restrict__ y) {
    if (MODE == MODE_SPMV) {
        y[i] = t;
    } else if (MODE == MODE_ADD) {
        y[i] = y[i] + t;
    } else if (MODE == MODE_RESID) {
        y[i] = b[i] - t;
    } else {
        TV r = b[i] - t;
        y[i] = x[i] + dval * r;
    }
}

// ---------------------------------------------------------------------------------------------
// upload helpers
// ---------------------------------------------------------------------------------------------
__global__ void convert_index_kernel(const long long* __restrict__ in, int* __restrict__ out, long long n,
                                     long long base) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = (int)(in[i] - base);
}

template <typename T>
__global__ void conj_copy_kernel(const T* __restrict__ in, T* __restrict__ out, long long n, int do_conj) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = do_conj ? conj_(in[i]) : in[i];
}

// max over row blocks of the 4-aligned nnz span, and max row length
__global__ void chunk_span_kernel(const int* __restrict__ rowptr, int n_rows, int rpc, int* __restrict__ out) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    int nchunks = (n_rows + rpc - 1) / rpc;
    if (c >= nchunks) return;
    int r0 = c * rpc;
    int r1 = min(r0 + rpc, n_rows);
    int ka = rowptr[r0] & ~3;
    int kb = (rowptr[r1] + 3) & ~3;
    atomicMax(out, kb - ka);
}

__global__ void max_rowlen_kernel(const int* __restrict__ rowptr, int n_rows, int* __restrict__ out) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    atomicMax(out, rowptr[r + 1] - rowptr[r]);
}

// ---------------------------------------------------------------------------------------------
// single right-hand side
// ---------------------------------------------------------------------------------------------
template <typename TA, typename TV, int TPR, int MODE>
__global__ void csr_stream_kernel(int n_rows, const int* __restrict__ rowptr, const int* __restrict__ colind,
                                  const TA* __restrict__ val, const TV* __restrict__ x,
                                  const TV* __restrict__ b, const TV* __restrict__ d, TV* __restrict__ y,
                                  int rows_per_cta, int cap) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    TA* sval = reinterpret_cast<TA*>(smem_raw + 16);
    int* scol = reinterpret_cast<int*>(smem_raw + 16 + (size_t)cap * sizeof(TA));

    const int tid = threadIdx.x;
    const int R0 = blockIdx.x * rows_per_cta;
    const int R1 = min(R0 + rows_per_cta, n_rows);
    const int k0 = __ldg(rowptr + R0);
    const int k1 = __ldg(rowptr + R1);
    const int ka = k0 & ~3;
    const int cnt = ((k1 + 3) & ~3) - ka;

    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0 && cnt > 0) {
        mbar_expect_tx(bar, (uint32_t)(cnt * (sizeof(TA) + sizeof(int))));
        bulk_g2s(sval, val + ka, (uint32_t)(cnt * sizeof(TA)), bar);
        bulk_g2s(scol, colind + ka, (uint32_t)(cnt * sizeof(int)), bar);
    }

    const int row = R0 + tid / TPR;
    const int g = tid % TPR;
    const bool active = row < R1;
    int rs = 0, re = 0;
    TV bval = VT<TV>::zero(), dval = VT<TV>::zero(), xval = VT<TV>::zero();
    if (active) {
        rs = __ldg(rowptr + row) - ka;
        re = __ldg(rowptr + row + 1) - ka;
        if (g == 0) {
            if (MODE == MODE_RESID || MODE == MODE_SWEEP) bval = b[row];
            if (MODE == MODE_SWEEP) {
                dval = d[row];
                xval = x[row];
            }
            if (MODE == MODE_ADD) xval = y[row];
        }
    }
    if (cnt > 0) mbar_wait(bar, 0);

    TV acc = VT<TV>::zero();
#pragma unroll 4
    for (int k = rs + g; k < re; k += TPR) {
        acc = acc + sval[k] * ldg_(x + scol[k]);
    }
    if (TPR > 1) {
#pragma unroll
        for (int s = TPR / 2; s > 0; s >>= 1) acc = acc + shfl_xor_(acc, s);
    }
    if (active && g == 0) {
        if (MODE == MODE_SPMV) {
            y[row] = acc;
        } else if (MODE == MODE_ADD) {
            y[row] = xval + acc;
        } else if (MODE == MODE_RESID) {
            y[row] = bval - acc;
        } else {
            TV r = bval - acc;
            y[row] = xval + dval * r;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// m right-hand sides, stored RHS-fastest: x[col*m + j].  One pass over the matrix serves all
// columns (the reference re-streams the matrix once per column).
// MP = lanes over right-hand sides (power of two <= 32); NT/MP rows are in flight per pass.
// ---------------------------------------------------------------------------------------------
template <typename TA, typename TV, int MODE>
__global__ void csr_stream_mrhs_kernel(int n_rows, const int* __restrict__ rowptr,
                                       const int* __restrict__ colind, const TA* __restrict__ val,
                                       const TV* __restrict__ x, const TV* __restrict__ b,
                                       const TV* __restrict__ d, TV* __restrict__ y, int rows_per_cta,
                                       int cap, int m, int mp) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    TA* sval = reinterpret_cast<TA*>(smem_raw + 16);
    int* scol = reinterpret_cast<int*>(smem_raw + 16 + (size_t)cap * sizeof(TA));

    const int tid = threadIdx.x;
    const int R0 = blockIdx.x * rows_per_cta;
    const int R1 = min(R0 + rows_per_cta, n_rows);
    const int k0 = __ldg(rowptr + R0);
    const int k1 = __ldg(rowptr + R1);
    const int ka = k0 & ~3;
    const int cnt = ((k1 + 3) & ~3) - ka;
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0 && cnt > 0) {
        mbar_expect_tx(bar, (uint32_t)(cnt * (sizeof(TA) + sizeof(int))));
        bulk_g2s(sval, val + ka, (uint32_t)(cnt * sizeof(TA)), bar);
        bulk_g2s(scol, colind + ka, (uint32_t)(cnt * sizeof(int)), bar);
    }
    const int j0 = tid % mp;
    const int rsub = tid / mp;
    const int rstep = blockDim.x / mp;
    if (cnt > 0) mbar_wait(bar, 0);
    for (int row = R0 + rsub; row < R1; row += rstep) {
        const int rs = __ldg(rowptr + row) - ka;
        const int re = __ldg(rowptr + row + 1) - ka;
        const TV dval = (MODE == MODE_SWEEP) ? d[row] : VT<TV>::zero();
        for (int j = j0; j < m; j += mp) {
            TV acc = VT<TV>::zero();
#pragma unroll 4
            for (int k = rs; k < re; ++k) acc = acc + sval[k] * ldg_(x + (long long)scol[k] * m + j);
            csr_epilogue<MODE, TV>(acc, (long long)row * m + j, dval, x, b, y);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// fallback for matrices whose row blocks do not fit shared memory (very long rows):
// one warp per row straight from global memory.
// ---------------------------------------------------------------------------------------------
template <typename TA, typename TV, int MODE>
__global__ void csr_rowwarp_kernel(int n_rows, const int* __restrict__ rowptr, const int* __restrict__ colind,
                                   const TA* __restrict__ val, const TV* __restrict__ x,
                                   const TV* __restrict__ b, const TV* __restrict__ d, TV* __restrict__ y,
                                   int m) {
    const int lane = threadIdx.x & 31;
    const int row = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5);
    if (row >= n_rows) return;
    const int rs = __ldg(rowptr + row), re = __ldg(rowptr + row + 1);
    const TV dval = (MODE == MODE_SWEEP) ? d[row] : VT<TV>::zero();
    for (int j = 0; j < m; ++j) {
        TV acc = VT<TV>::zero();
        for (int k = rs + lane; k < re; k += 32) acc = acc + ldg_(val + k) * ldg_(x + (long long)ldg_(colind + k) * m + j);
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) acc = acc + shfl_xor_(acc, s);
        if (lane == 0) csr_epilogue<MODE, TV>(acc, (long long)row * m + j, dval, x, b, y);
    }
}

__device__ __forceinline__ double ldg_val(const double* p) { return __ldg(p); }

}  // namespace mgb200
