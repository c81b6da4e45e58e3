// Device-side re-setup of a hierarchy whose fine matrix changed: replaceMatrixInHierarchy of the reference
// (src/Multigrid/MGsetup.jl:226-270) keeps Ps / Rs and redoes, per level, the relaxation weights
// (getRelaxPrec, MGsetup.jl:142-160, :359-362), the Galerkin product `Ps[l]*AT*Rs[l]` (:259) and the coarsest
// factorisation (defineCoarsestAinv, :323-355).  P and R are fixed, so the SPARSITY of every coarse operator is the one
// the first setup produced and is already on the device as the CSR structure of A_{l+1}: the product below is purely
// numeric - no symbolic phase, no hashing, no product plan (a plan of index triples for the 27-point Galerkin levels
// would be ~150 triples per non-zero, larger than the hierarchy).
//
//   A_c[i, j] = sum_k R[i, k] * ( sum_l A[k, l] * P[l, j] )        k, l in stored order
//
// One warp owns coarse row i; lane s owns slot s of that row (column j = colind_c[rowptr_c[i] + s], 32 slots per pass).
// All lanes walk the same (k, l, q) sequence - every load of the walk is warp-uniform, i.e. one broadcast - and a lane
// takes the product when P's column matches its own.  Each output entry is therefore accumulated by ONE lane in a fixed
// order: the result is deterministic (no atomics), and equals the row-by-row (Gustavson) evaluation of T = A P, A_c = R T.
// Work: |R row| * |A row| * |P row| steps per pass, 27 * 7 * 3.4 = 640 for the 7-point fine level of cfg2 - 2.1 M coarse rows
// in a few milliseconds, against seconds for the host product + re-upload.
#pragma once
#include "pattern.cuh"

namespace mgb200 {

template <typename TV, typename RT>
__global__ void __launch_bounds__(256) galerkin_kernel(int nc, const int* __restrict__ r_ptr, const int* __restrict__ r_col,
                                                       const RT* __restrict__ r_val, const int* __restrict__ a_ptr,
                                                       const int* __restrict__ a_col, const TV* __restrict__ a_val,
                                                       const int* __restrict__ p_ptr, const int* __restrict__ p_col,
                                                       const RT* __restrict__ p_val, const int* __restrict__ c_ptr,
                                                       const int* __restrict__ c_col, TV* __restrict__ c_val) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < nc; i += gridDim.x * wpb) {
        const int c0 = c_ptr[i], c1 = c_ptr[i + 1];
        const int k0 = r_ptr[i], k1 = r_ptr[i + 1];
        for (int s0 = c0; s0 < c1; s0 += 32) {
            const int s = s0 + lane;
            const int mycol = s < c1 ? c_col[s] : -1;
            TV acc = VT<TV>::zero();
            for (int kk = k0; kk < k1; ++kk) {
                const int k = r_col[kk];
                const RT rv = r_val[kk];
                TV t = VT<TV>::zero();
                bool any = false;
                const int l0 = a_ptr[k], l1 = a_ptr[k + 1];
                for (int ll = l0; ll < l1; ++ll) {
                    const int l = a_col[ll];
                    const TV av = a_val[ll];
                    const int q0 = p_ptr[l], q1 = p_ptr[l + 1];
                    for (int q = q0; q < q1; ++q)
                        if (p_col[q] == mycol) {
                            t = t + av * p_val[q];
                            any = true;
                        }
                }
                if (any) acc = acc + rv * t;
            }
            if (s < c1) c_val[s] = acc;
        }
    }
}

// same sparsity?  (row pointers and column indices of the new fine matrix against the resident one)
__global__ void csr_same_structure_kernel(long long n, const int* __restrict__ a, const int* __restrict__ b, int* __restrict__ flag) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        if (a[i] != b[i]) *flag = 1;
}

// getRelaxPrec (MGsetup.jl:142-160) from the resident CSR of the OPERATOR (the stored arrays are those of A^H, conjugated
// at upload).  kind 0, "Jac": d = conj(omega ./ diag(AT)) = omega / a_ii.  kind 1, "SPAI" (getSPAIprec, :359-362):
// d = conj(omega * conj(diag(AT)) ./ rowsumsq(AT)) = omega * conj(a_ii) / ||A e_i||^2 - the COLUMN i of A, which a
// structurally symmetric CSR reaches through the rows its own row points at (entry (j, i) of every j in row i; summed in
// the stored order of row i, so the result is deterministic).  flag bit 0: a row has no diagonal entry; bit 1: the
// structure is not symmetric (the caller falls back to the host setup).
template <typename TV>
__global__ void relax_prec_kernel(int n, const int* __restrict__ ptr, const int* __restrict__ col, const TV* __restrict__ val,
                                  int kind, double omega, TV* __restrict__ d, int* __restrict__ flag) {
    typedef typename VT<TV>::real_t R;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        TV aii = VT<TV>::zero();
        bool have = false;
        double ss = 0.0;
        for (int k = ptr[i]; k < ptr[i + 1]; ++k) {
            const int j = col[k];
            if (j == i) {
                aii = val[k];
                have = true;
            }
            if (kind == 1) {
                bool found = false;
                for (int q = ptr[j]; q < ptr[j + 1]; ++q)
                    if (col[q] == i) {
                        const double re = VT<TV>::re(val[q]), im = VT<TV>::im(val[q]);
                        ss = ss + (re * re + im * im);
                        found = true;
                        break;
                    }
                if (!found) atomicOr(flag, 2);
            }
        }
        if (!have) atomicOr(flag, 1);
        const double re = VT<TV>::re(aii), im = VT<TV>::im(aii);
        if (kind == 0) {
            // omega / a_ii; complex: omega * conj(a_ii) / |a_ii|^2
            const double den = VT<TV>::is_complex ? (re * re + im * im) : re;
            d[i] = VT<TV>::is_complex ? VT<TV>::make((R)(omega * re / den), (R)(-(omega * im) / den)) : VT<TV>::make((R)(omega / den), 0.0);
        } else {
            d[i] = VT<TV>::make((R)(omega * re / ss), (R)(-(omega * im) / ss));
        }
    }
}

// ---- stencil dictionary of a matrix whose values changed (pattern.cuh) -----------------------------------------------------
// The dictionary stays valid when every row still carries the values of its pattern's representative row, bit for bit.
template <typename TA>
__global__ void pat_verify_values_kernel(int n, const int* __restrict__ ptr, const TA* __restrict__ val,
                                         const uint16_t* __restrict__ pid, const int* __restrict__ rep, int* __restrict__ flag) {
    constexpr int W = sizeof(TA) / 4;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int r = rep[pid[i]];
        if (r == i) continue;
        const int a = ptr[i], len = ptr[i + 1] - a, b = ptr[r];
        if (len != ptr[r + 1] - b) {
            *flag = 1;
            continue;
        }
        const unsigned* x = reinterpret_cast<const unsigned*>(val + a);
        const unsigned* y = reinterpret_cast<const unsigned*>(val + b);
        bool same = true;
        for (int k = 0; k < len * W; ++k) same = same && (x[k] == y[k]);
        if (!same) *flag = 1;
    }
}
template <typename TA>
__global__ void pat_refresh_values_kernel(int npat, const int* __restrict__ pat_off, const int* __restrict__ rep,
                                          const int* __restrict__ ptr, const TA* __restrict__ val, PatEntry<TA>* __restrict__ ent,
                                          PatEntry<TA>* __restrict__ ent_s, TA* __restrict__ vals_out) {
    const int p = blockIdx.x;
    if (p >= npat) return;
    const int e0 = pat_off[p], len = pat_off[p + 1] - e0, a = ptr[rep[p]];
    for (int k = threadIdx.x; k < len; k += blockDim.x) {
        const TA v = val[a + k];
        ent[e0 + k].v = v;
        if (ent_s) ent_s[e0 + k].v = v;
        vals_out[e0 + k] = v;
    }
}
// d as a function of the pattern id: dp[p] = d[rep[p]], valid when every row agrees bit for bit
template <typename TV>
__global__ void pat_gather_d_kernel(int npat, const int* __restrict__ rep, const TV* __restrict__ d, TV* __restrict__ dp) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < npat) dp[p] = d[rep[p]];
}
template <typename TV>
__global__ void pat_verify_d_kernel(int n, const uint16_t* __restrict__ pid, const TV* __restrict__ dp, const TV* __restrict__ d,
                                    int* __restrict__ flag) {
    constexpr int W = sizeof(TV) / 4;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned* x = reinterpret_cast<const unsigned*>(d + i);
        const unsigned* y = reinterpret_cast<const unsigned*>(dp + pid[i]);
        bool same = true;
        for (int k = 0; k < W; ++k) same = same && (x[k] == y[k]);
        if (!same) *flag = 1;
    }
}

}  // namespace mgb200
