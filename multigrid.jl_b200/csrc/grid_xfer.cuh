// Transfer operators of the geometric hierarchy with a GRID HINT (off by default: option "grid_transfers").
//
// P = kron of 1-D linear interpolations on nodal grids, R = 2^-dim P^T (src/Multigrid/GeometricTransferOperators.jl:5-46,
// MGsetup.jl:53-62 of the reference).  In stencil-dictionary form (pattern.cuh) their rows are copies of 8 (P: one per
// parity class of the fine node) or 27 (R: first / interior / last per dimension) patterns, but the dictionary kernel
// still reads 6 bytes per row (pattern id + first column) and walks the entries per lane.  The reference's setup knows
// the meshes (param.Meshes[l].n); when the host passes them (mgb200_set_level_grid) and the uploaded P / R are exactly
// what those grids imply - verified row by row at upload, values untouched - the pattern and the columns of a row are
// functions of its grid coordinates: the kernels below read NO matrix stream at all, only the dictionary VALUES (so a
// P with other weights still works), and block R coarse lines per thread:
//   restriction:  a thread owns coarse column I of R coarse lines; per fine plane it loads the 3 x (2R+1) fine values
//                 around its coarse nodes once for 27 R products;
//   prolongation: it loads the 2 x (R+1) x 2 coarse corner values once and updates the 8 R fine nodes of its cells.
// Entries are multiplied in stored order ((dz,dy,dx) ascending), so results are bit-identical to the dictionary walk.
// The per-thread functions are __host__ __device__: mgb200_host_grid_transfer runs them on the CPU
// (tests/test_patterns.py).  NOT yet run on a GPU (written after the GPU budget of round 1 was spent).
#pragma once
#include "pattern.cuh"

namespace mgb200 {

struct GridXfer {
    int ok;            // 0: no hint / hint does not match the matrix
    int kind;          // 1: prolongation (rows = fine nodes), 2: restriction (rows = coarse nodes)
    int n[3], N[3];    // fine / coarse nodes per dimension (unused dimensions: 1); n = 2N - 1 where N > 1
    int cls_k0[27];    // first dictionary entry of the pattern of each class (-1: class does not occur)
};
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
static bool gx_verify_prolongation(const HostPatterns<TA>& H, long long n_rows, const int n[3], const int N[3], GridXfer& X) {
    X = no_grid();
    if (!H.ok || H.rowrel) return false;
    for (int d = 0; d < 3; ++d) {
        if (N[d] < 1 || n[d] != (N[d] > 1 ? 2 * N[d] - 1 : 1)) return false;
        X.n[d] = n[d];
        X.N[d] = N[d];
    }
    if ((long long)n[0] * n[1] * n[2] != n_rows) return false;
    int cls_pat[8];
    for (int c = 0; c < 8; ++c) cls_pat[c] = -1;
    for (int c = 0; c < 27; ++c) X.cls_k0[c] = -1;
    const long long N1 = N[0], N12 = (long long)N[0] * N[1];
    long long row = 0;
    for (int k = 0; k < n[2]; ++k)
        for (int j = 0; j < n[1]; ++j)
            for (int i = 0; i < n[0]; ++i, ++row) {
                const int a = i & 1, b = j & 1, c = k & 1, cls = a + 2 * b + 4 * c, p = H.pid[row];
                if (H.c0[row] != (i >> 1) + N1 * (j >> 1) + N12 * (k >> 1)) return false;
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
static bool gx_verify_restriction(const HostPatterns<TA>& H, long long n_rows, const int n[3], const int N[3], GridXfer& X) {
    X = no_grid();
    if (!H.ok || H.rowrel) return false;
    for (int d = 0; d < 3; ++d) {
        if (N[d] < 1 || n[d] != (N[d] > 1 ? 2 * N[d] - 1 : 1)) return false;
        X.n[d] = n[d];
        X.N[d] = N[d];
    }
    if ((long long)N[0] * N[1] * N[2] != n_rows) return false;
    int cls_pat[27];
    for (int c = 0; c < 27; ++c) {
        cls_pat[c] = -1;
        X.cls_k0[c] = -1;
    }
    const long long S = n[0], S2 = (long long)n[0] * n[1];
    long long row = 0;
    for (int K = 0; K < N[2]; ++K)
        for (int J = 0; J < N[1]; ++J)
            for (int I = 0; I < N[0]; ++I, ++row) {
                const int cx = gx_class(I, N[0]), cy = gx_class(J, N[1]), cz = gx_class(K, N[2]), cls = cx + 3 * cy + 9 * cz;
                const int ax = gx_allowed(cx, N[0]), ay = gx_allowed(cy, N[1]), az = gx_allowed(cz, N[2]);
                const int p = H.pid[row];
                const long long anchor = 2LL * I + S * (2LL * J) + S2 * (2LL * K);
                // first entry = smallest allowed offset in every dimension
                const int fx = (ax & 1) ? -1 : 0, fy = (ay & 1) ? -1 : 0, fz = (az & 1) ? -1 : 0;
                const long long first = fx + S * fy + S2 * fz;
                if (H.c0[row] != anchor + first) return false;
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

// ---- restriction: r_c = R r_f, thread = coarse column I of the coarse lines [J0, J0+R) of coarse plane K ------------------
template <typename TA, typename TV, int R>
__host__ __device__ inline void gx_restrict_thread(const GridXfer& X, int I, int J0, int K, const PatEntry<TA>* ent,
                                                   const TV* rf, TV* rc) {
    const int N1 = X.N[0], N2 = X.N[1];
    const int nr = (N2 - J0 < R) ? (N2 - J0) : R;
    // one pattern for the R rows: all of them interior lines (or R == 1)
    const bool uniform = (R == 1) ? (nr == 1) : (nr == R && J0 >= 1 && J0 + R - 1 <= N2 - 2);
    if (!uniform) {
        for (int j = 0; j < nr; ++j) gx_restrict_thread<TA, TV, 1>(X, I, J0 + j, K, ent, rf, rc);
        return;
    }
    const int cx = gx_class(I, N1), cy = (R == 1) ? gx_class(J0, N2) : 1, cz = gx_class(K, X.N[2]);
    const int ax = gx_allowed(cx, N1), ay = gx_allowed(cy, N2), az = gx_allowed(cz, X.N[2]);
    const PatEntry<TA>* e = ent + X.cls_k0[cx + 3 * cy + 9 * cz];
    const long long S = X.n[0], S2 = (long long)X.n[0] * X.n[1];
    const long long f0 = S2 * (2LL * K) + S * (2LL * J0) + 2LL * I;       // fine node under the first coarse node
    const long long crow0 = ((long long)K * N2 + J0) * N1 + I;
    TV acc[R];
#pragma unroll
    for (int j = 0; j < R; ++j) acc[j] = VT<TV>::zero();
#pragma unroll
    for (int dz = -1; dz <= 1; ++dz) {
        if (!((az >> (dz + 1)) & 1)) continue;
        const TV* xp = rf + (f0 + dz * S2 - S);                            // fine line 2 J0 - 1
        TV Xv[3][2 * R + 1];                                               // l = fine line - (2 J0 - 1)
#pragma unroll
        for (int l = 0; l < 2 * R + 1; ++l) {
            // coarse row j multiplies fine line l = 2j + 1 + dy: a value is loaded iff some row multiplies it
            const bool lneed = (l & 1) ? ((ay >> 1) & 1) : (((l <= 2 * R - 2) && (ay & 1)) || ((l >= 2) && ((ay >> 2) & 1)));
            const TV* q = xp + l * S;
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) Xv[dx + 1][l] = (lneed && ((ax >> (dx + 1)) & 1)) ? ld_ro(q + dx) : VT<TV>::zero();
        }
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
                if (((ay >> (dy + 1)) & 1) && ((ax >> (dx + 1)) & 1)) {
                    const TA v = e->v;
                    ++e;
#pragma unroll
                    for (int j = 0; j < R; ++j) acc[j] = acc[j] + v * Xv[dx + 1][2 * j + 1 + dy];
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < R; ++j) rc[crow0 + (long long)j * N1] = acc[j];
}

// ---- prolongation: x_f += P x_c, thread = coarse column I of the coarse lines [J0, J0+R) of coarse plane K; it owns the
// fine nodes (2I+a, 2J+b, 2K+c), a, b, c in {0,1}, of those coarse nodes -------------------------------------------------
template <typename TA, typename TV, int R>
__host__ __device__ inline void gx_prolong_thread(const GridXfer& X, int I, int J0, int K, const PatEntry<TA>* ent,
                                                  const TV* xc, TV* xf) {
    const int N1 = X.N[0], N2 = X.N[1], N3 = X.N[2];
    const long long S = X.n[0], S2 = (long long)X.n[0] * X.n[1], CS = N1, CS2 = (long long)N1 * N2;
    const long long crow0 = ((long long)K * N2 + J0) * N1 + I;
    const bool hasI = I + 1 < N1, hasK = K + 1 < N3;
    TV XC[2][R + 1][2];
#pragma unroll
    for (int dz = 0; dz < 2; ++dz)
#pragma unroll
        for (int l = 0; l < R + 1; ++l) {
            const bool in = (J0 + l < N2) && (dz == 0 || hasK);
            const TV* q = xc + (crow0 + dz * CS2 + l * CS);
            XC[dz][l][0] = in ? ld_ro(q) : VT<TV>::zero();
            XC[dz][l][1] = (in && hasI) ? ld_ro(q + 1) : VT<TV>::zero();
        }
#pragma unroll
    for (int c = 0; c < 2; ++c) {
#pragma unroll
        for (int b = 0; b < 2; ++b) {
#pragma unroll
            for (int a = 0; a < 2; ++a) {
                if ((c == 0 || hasK) && (a == 0 || hasI)) {
                    const int k0 = X.cls_k0[a + 2 * b + 4 * c];
                    if (k0 >= 0) {
                        // the values of this parity class in stored order (dz', dy', dx'): 2^(a+b+c) of them
                        TA v[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) v[k] = (k < (1 << (a + b + c))) ? ent[k0 + k].v : TA();
#pragma unroll
                        for (int j = 0; j < R; ++j) {
                            if (J0 + j < N2 && (b == 0 || J0 + j + 1 < N2)) {
                                const long long row = S2 * (2LL * K + c) + S * (2LL * (J0 + j) + b) + 2LL * I + a;
                                TV acc = VT<TV>::zero();
#pragma unroll
                                for (int dz = 0; dz <= c; ++dz)
#pragma unroll
                                    for (int dy = 0; dy <= b; ++dy)
#pragma unroll
                                        for (int dx = 0; dx <= a; ++dx)
                                            acc = acc + v[(dz * (b + 1) + dy) * (a + 1) + dx] * XC[dz][j + dy][dx];
                                xf[row] = xf[row] + acc;
                            }
                        }
                    }
                }
            }
        }
    }
}

// persistent grid-stride over the flattened (plane, line group, column) index of the COARSE grid
template <typename TA, typename TV, int KIND, int R>
__global__ void __launch_bounds__(256) gx_kernel(const __grid_constant__ GridXfer X, long long total,
                                                 const PatEntry<TA>* __restrict__ ent, const TV* __restrict__ in,
                                                 TV* __restrict__ out) {
    const int N1 = X.N[0], gpp = (X.N[1] + R - 1) / R;
    for (long long f = blockIdx.x * (long long)blockDim.x + threadIdx.x; f < total; f += (long long)gridDim.x * blockDim.x) {
        const long long grp = f / N1;
        const int I = (int)(f - grp * N1), K = (int)(grp / gpp), q = (int)(grp - (long long)K * gpp);
        if (KIND == 2) gx_restrict_thread<TA, TV, R>(X, I, q * R, K, ent, in, out);
        else gx_prolong_thread<TA, TV, R>(X, I, q * R, K, ent, in, out);
    }
}

}  // namespace mgb200
