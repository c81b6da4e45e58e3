// Micro-benchmark for line-blocked forms of the TRANSFER operators in stencil-dictionary form (DESIGN.md section 9,
// item 2): restriction r_c = R r_f and prolongation x_f += P x_c of the geometric hierarchy (P = kron of 1-D linear
// interpolations on nodal grids, R = 2^-3 P^T; src/Multigrid/GeometricTransferOperators.jl:5-46, MGsetup.jl:53-62 of
// the reference).
//
// The production kernel (csrc/pattern.cuh, pat_kernel with offsets from the first stored column) gives one row to one
// thread: per entry one 128-bit dictionary load (four quarter-warp passes even when every lane reads the same entry)
// and one gather whose lanes are 16 bytes apart.  Here a thread owns a coarse column I of R consecutive coarse lines:
//   restriction:  per fine plane 2K+dz it loads the 3 x (2R+1) fine values around its coarse nodes once and every
//                 dictionary value once for R coarse rows (27 R products from 9 (2R+1) loads instead of 27 R);
//   prolongation: it loads the 2 x (R+1) x 2 coarse corner values once and produces the 8R fine rows of its cells.
// A pattern is a presence mask plus its values in stored order; stored order is (dz,dy,dx) order, so every row still
// accumulates in stored order: bit-identical to the one-row-per-thread kernels.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -lineinfo -o gpurun_out/microbench_transfer tools/microbench_transfer.cu
//   gpurun_out/microbench_transfer               # 129^3 -> 257^3 on the GPU
//   gpurun_out/microbench_transfer --host-check  # no GPU: the per-thread functions on the CPU for small grids
//
// NOT part of the library; not yet run on a B200 (the GPU budget of the round was spent); the host check passes.
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

struct __align__(16) Ent { double v; int delta; int pad; };
struct Pat { int k0, len, mask; };
// coarse grid N1 x N2 x N3 nodes, fine grid n = 2N - 1 per dimension
struct Geo { int N1, N2, N3, n1, n2, n3; long long NC, nf; };

static Geo make_geo(int N1, int N2, int N3) {
    Geo g; g.N1 = N1; g.N2 = N2; g.N3 = N3; g.n1 = 2 * N1 - 1; g.n2 = 2 * N2 - 1; g.n3 = 2 * N3 - 1;
    g.NC = (long long)N1 * N2 * N3; g.nf = (long long)g.n1 * g.n2 * g.n3;
    return g;
}

// ---- reference: one row per thread, column = c0[row] + delta (what pat_kernel<..., ROWREL = false> does) ----------
__host__ __device__ inline double dict_row(long long row, const uint16_t* pid, const int* c0, const Pat* pat, const Ent* ent,
                                           const double* x) {
    const Pat P = pat[pid[row]];
    const long long base = c0[row];
    double acc = 0.0;
    for (int k = P.k0; k < P.k0 + P.len; ++k) acc = acc + ent[k].v * x[base + ent[k].delta];
    return acc;
}
__global__ void __launch_bounds__(256) k_ref(long long n, int add, const uint16_t* __restrict__ pid, const int* __restrict__ c0,
                                             const Pat* __restrict__ pat, const Ent* __restrict__ ent,
                                             const double* __restrict__ x, double* __restrict__ y) {
    const long long row = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (row >= n) return;
    const double acc = dict_row(row, pid, c0, pat, ent, x);
    y[row] = add ? y[row] + acc : acc;
}

// ---- restriction: thread = coarse column I of the coarse lines [J0, J0+R) of coarse plane K ------------------------
template <int R>
__host__ __device__ inline void restrict_thread(const Geo& g, int I, int J0, int K, const uint16_t* pid, const int* c0,
                                                const Pat* pat, const Ent* ent, const double* rf, double* rc) {
    const long long crow0 = ((long long)K * g.N2 + J0) * g.N1 + I;
    const int nr = (g.N2 - J0 < R) ? (g.N2 - J0) : R;
    const int p0 = pid[crow0];
    bool same = (nr == R);
#pragma unroll
    for (int j = 1; j < R; ++j)
        if (j < nr) same = same && (pid[crow0 + (long long)j * g.N1] == p0);
    if (!same) {
        for (int j = 0; j < nr; ++j) rc[crow0 + (long long)j * g.N1] = dict_row(crow0 + (long long)j * g.N1, pid, c0, pat, ent, rf);
        return;
    }
    const Pat P = pat[p0];
    const Ent* e = ent + P.k0;
    const long long S = g.n1, S2 = (long long)g.n1 * g.n2;
    const long long f0 = (long long)(2 * K) * S2 + (long long)(2 * J0) * S + 2 * I;    // fine node under the first coarse node
    double acc[R];
#pragma unroll
    for (int j = 0; j < R; ++j) acc[j] = 0.0;
#pragma unroll
    for (int dz = -1; dz <= 1; ++dz) {
        const int mz = (P.mask >> ((dz + 1) * 9)) & 0x1FF;
        if (mz == 0) continue;
        const double* xp = rf + f0 + (long long)dz * S2 - S;        // fine line 2 J0 - 1
        double X[3][2 * R + 1];                                     // l = fine line - (2 J0 - 1)
        // coarse row j uses fine line l = 2j + 1 + dy: a value is loaded iff some row multiplies it, so every load is an
        // address some row's entry names
#pragma unroll
        for (int l = 0; l < 2 * R + 1; ++l) {
            const double* q = xp + (long long)l * S;
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
                const bool need = (l & 1) ? (mz & (1 << (3 + dx + 1))) != 0
                                          : (((l <= 2 * R - 2) && (mz & (1 << (0 + dx + 1)))) || ((l >= 2) && (mz & (1 << (6 + dx + 1)))));
                X[dx + 1][l] = need ? q[dx] : 0.0;
            }
        }
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
                if (mz & (1 << ((dy + 1) * 3 + (dx + 1)))) {
                    const double v = e->v;
                    ++e;
#pragma unroll
                    for (int j = 0; j < R; ++j) acc[j] = acc[j] + v * X[dx + 1][2 * j + 1 + dy];
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < R; ++j) rc[crow0 + (long long)j * g.N1] = acc[j];
}
template <int R>
__global__ void __launch_bounds__(256) k_restrict_lines(Geo g, const uint16_t* __restrict__ pid, const int* __restrict__ c0,
                                                        const Pat* __restrict__ pat, const Ent* __restrict__ ent,
                                                        const double* __restrict__ rf, double* __restrict__ rc) {
    const int gpp = (g.N2 + R - 1) / R;
    const long long total = (long long)g.N3 * gpp * g.N1;
    for (long long f = blockIdx.x * (long long)blockDim.x + threadIdx.x; f < total; f += (long long)gridDim.x * blockDim.x) {
        const long long grp = f / g.N1;
        const int I = (int)(f - grp * g.N1), K = (int)(grp / gpp), q = (int)(grp - (long long)K * gpp);
        restrict_thread<R>(g, I, q * R, K, pid, c0, pat, ent, rf, rc);
    }
}

// ---- prolongation: thread = coarse column I of the coarse lines [J0, J0+R) of coarse plane K; it owns the fine nodes
// (2I+a, 2J+b, 2K+c), a, b, c in {0,1}, of those coarse nodes.  The pattern of a fine row is its parity class a+2b+4c
// (checked against pid; anything else falls back to the dictionary walk).
template <int R>
__host__ __device__ inline void prolong_thread(const Geo& g, int I, int J0, int K, const uint16_t* pid, const int* c0,
                                               const Pat* pat, const Ent* ent, const double* xc, double* xf) {
    const long long S = g.n1, S2 = (long long)g.n1 * g.n2, CS = g.N1, CS2 = (long long)g.N1 * g.N2;
    const long long crow0 = ((long long)K * g.N2 + J0) * g.N1 + I;
    const bool hasI = I + 1 < g.N1, hasK = K + 1 < g.N3;
    double XC[2][R + 1][2];
#pragma unroll
    for (int dz = 0; dz < 2; ++dz)
#pragma unroll
        for (int l = 0; l < R + 1; ++l) {
            const bool in = (J0 + l < g.N2) && (dz == 0 || hasK);
            const double* q = xc + crow0 + (long long)dz * CS2 + (long long)l * CS;
            XC[dz][l][0] = in ? q[0] : 0.0;
            XC[dz][l][1] = (in && hasI) ? q[1] : 0.0;
        }
#pragma unroll
    for (int c = 0; c < 2; ++c) {
#pragma unroll
        for (int bb = 0; bb < 2; ++bb) {
#pragma unroll
            for (int a = 0; a < 2; ++a) {
                if ((c == 0 || hasK) && (a == 0 || hasI)) {
                    const int cls = a + 2 * bb + 4 * c;
                    const Pat P = pat[cls];
                    const bool regular = (P.len == (1 << (a + bb + c)));
                    // the values of this parity class in stored order (dz', dy', dx'): 2^(a+b+c) of them
                    double v[8];
#pragma unroll
                    for (int k = 0; k < (1 << (a + bb + c)); ++k) v[k] = regular ? ent[P.k0 + k].v : 0.0;
#pragma unroll
                    for (int j = 0; j < R; ++j) {
                        if (J0 + j < g.N2 && (bb == 0 || J0 + j + 1 < g.N2)) {
                            const long long row = (long long)(2 * K + c) * S2 + (long long)(2 * (J0 + j) + bb) * S + 2 * I + a;
                            if (pid[row] != cls || !regular) {                  // not the expected interpolation row
                                xf[row] = xf[row] + dict_row(row, pid, c0, pat, ent, xc);
                            } else {
                                double acc = 0.0;
#pragma unroll
                                for (int dz = 0; dz <= c; ++dz)
#pragma unroll
                                    for (int dy = 0; dy <= bb; ++dy)
#pragma unroll
                                        for (int dx = 0; dx <= a; ++dx)
                                            acc = acc + v[(dz * (bb + 1) + dy) * (a + 1) + dx] * XC[dz][j + dy][dx];
                                xf[row] = xf[row] + acc;
                            }
                        }
                    }
                }
            }
        }
    }
}
template <int R>
__global__ void __launch_bounds__(256) k_prolong_lines(Geo g, const uint16_t* __restrict__ pid, const int* __restrict__ c0,
                                                       const Pat* __restrict__ pat, const Ent* __restrict__ ent,
                                                       const double* __restrict__ xc, double* __restrict__ xf) {
    const int gpp = (g.N2 + R - 1) / R;
    const long long total = (long long)g.N3 * gpp * g.N1;
    for (long long f = blockIdx.x * (long long)blockDim.x + threadIdx.x; f < total; f += (long long)gridDim.x * blockDim.x) {
        const long long grp = f / g.N1;
        const int I = (int)(f - grp * g.N1), K = (int)(grp / gpp), q = (int)(grp - (long long)K * gpp);
        prolong_thread<R>(g, I, q * R, K, pid, c0, pat, ent, xc, xf);
    }
}

// ---- the two operators in dictionary form ----------------------------------------------------------------------------
struct Dict { std::vector<uint16_t> pid; std::vector<int> c0; std::vector<Pat> pat; std::vector<Ent> ent; };

// P: fine rows, 8 patterns (parity classes), columns relative to the coarse node (I,J,K) = (i/2, j/2, k/2)
static void build_P(const Geo& g, Dict& D) {
    D.pid.resize(g.nf); D.c0.resize(g.nf); D.pat.clear(); D.ent.clear();
    for (int cls = 0; cls < 8; ++cls) {
        const int a = cls & 1, b = (cls >> 1) & 1, c = cls >> 2;
        Pat P; P.k0 = (int)D.ent.size(); P.mask = 0;
        for (int dz = 0; dz <= c; ++dz)
            for (int dy = 0; dy <= b; ++dy)
                for (int dx = 0; dx <= a; ++dx) {
                    Ent e; e.v = 1.0 / (1 << (a + b + c)); e.delta = dx + g.N1 * dy + g.N1 * g.N2 * dz; e.pad = 0;
                    D.ent.push_back(e);
                }
        P.len = (int)D.ent.size() - P.k0;
        D.pat.push_back(P);
    }
    for (int k = 0; k < g.n3; ++k)
        for (int j = 0; j < g.n2; ++j)
            for (int i = 0; i < g.n1; ++i) {
                const long long row = ((long long)k * g.n2 + j) * g.n1 + i;
                D.pid[row] = (uint16_t)((i & 1) + 2 * (j & 1) + 4 * (k & 1));
                D.c0[row] = (int)(((long long)(k >> 1) * g.N2 + (j >> 1)) * g.N1 + (i >> 1));
            }
}
// R = P^T / 8: coarse rows, 27 patterns (first / interior / last per dimension), columns relative to the first stored one
static void build_R(const Geo& g, Dict& D) {
    D.pid.resize(g.NC); D.c0.resize(g.NC); D.pat.clear(); D.ent.clear();
    std::vector<int> first(27);
    for (int cl = 0; cl < 27; ++cl) {
        const int cx = cl % 3, cy = (cl / 3) % 3, cz = cl / 9;
        Pat P; P.k0 = (int)D.ent.size(); P.mask = 0;
        bool have = false;
        int d0 = 0;
        for (int dz = -1; dz <= 1; ++dz)
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                    if ((cx == 0 && dx < 0) || (cx == 2 && dx > 0) || (cy == 0 && dy < 0) || (cy == 2 && dy > 0) ||
                        (cz == 0 && dz < 0) || (cz == 2 && dz > 0)) continue;
                    const int d = dx + g.n1 * dy + g.n1 * g.n2 * dz;
                    if (!have) { d0 = d; have = true; }
                    Ent e; e.v = 0.125 / (1 << ((dx != 0) + (dy != 0) + (dz != 0))); e.delta = d - d0; e.pad = 0;
                    D.ent.push_back(e);
                    P.mask |= 1 << ((dz + 1) * 9 + (dy + 1) * 3 + (dx + 1));
                }
        P.len = (int)D.ent.size() - P.k0;
        D.pat.push_back(P);
        first[cl] = d0;
    }
    auto cls1 = [](int i, int n) { return i == 0 ? 0 : (i == n - 1 ? 2 : 1); };
    for (int K = 0; K < g.N3; ++K)
        for (int J = 0; J < g.N2; ++J)
            for (int I = 0; I < g.N1; ++I) {
                const long long row = ((long long)K * g.N2 + J) * g.N1 + I;
                const int cl = cls1(I, g.N1) + 3 * cls1(J, g.N2) + 9 * cls1(K, g.N3);
                D.pid[row] = (uint16_t)cl;
                D.c0[row] = (int)(((long long)(2 * K) * g.n2 + 2 * J) * g.n1 + 2 * I + first[cl]);
            }
}

template <int R>
static bool host_check_one(const Geo& g) {
    Dict DP, DR;
    build_P(g, DP); build_R(g, DR);
    const long long pad = (long long)g.n1 * g.n2 + g.n1 + 2;
    std::vector<double> fbuf(g.nf + 2 * pad, 1e300), cbuf(g.NC + 2 * pad, 1e300);
    double *rf = fbuf.data() + pad, *xc = cbuf.data() + pad;
    srand(11);
    for (long long i = 0; i < g.nf; ++i) rf[i] = rand() / (double)RAND_MAX;
    for (long long i = 0; i < g.NC; ++i) xc[i] = rand() / (double)RAND_MAX;
    // restriction
    std::vector<double> rc_ref(g.NC), rc(g.NC, -1.0);
    for (long long r = 0; r < g.NC; ++r) rc_ref[r] = dict_row(r, DR.pid.data(), DR.c0.data(), DR.pat.data(), DR.ent.data(), rf);
    const int gpp = (g.N2 + R - 1) / R;
    for (int K = 0; K < g.N3; ++K)
        for (int q = 0; q < gpp; ++q)
            for (int I = 0; I < g.N1; ++I)
                restrict_thread<R>(g, I, q * R, K, DR.pid.data(), DR.c0.data(), DR.pat.data(), DR.ent.data(), rf, rc.data());
    const bool okR = memcmp(rc.data(), rc_ref.data(), g.NC * sizeof(double)) == 0;
    // prolongation (x_f += P x_c)
    std::vector<double> xf_ref(rf, rf + g.nf), xf(rf, rf + g.nf);
    for (long long r = 0; r < g.nf; ++r) xf_ref[r] = xf_ref[r] + dict_row(r, DP.pid.data(), DP.c0.data(), DP.pat.data(), DP.ent.data(), xc);
    for (int K = 0; K < g.N3; ++K)
        for (int q = 0; q < gpp; ++q)
            for (int I = 0; I < g.N1; ++I)
                prolong_thread<R>(g, I, q * R, K, DP.pid.data(), DP.c0.data(), DP.pat.data(), DP.ent.data(), xc, xf.data());
    const bool okP = memcmp(xf.data(), xf_ref.data(), g.nf * sizeof(double)) == 0;
    printf("host check %dx%dx%d -> %dx%dx%d R=%d: restriction %s, prolongation %s\n", g.N1, g.N2, g.N3, g.n1, g.n2, g.n3, R,
           okR ? "bit-identical" : "MISMATCH", okP ? "bit-identical" : "MISMATCH");
    return okR && okP;
}
static int host_check() {
    bool ok = true;
    const int grids[4][3] = {{5, 5, 5}, {9, 6, 4}, {3, 7, 2}, {17, 3, 3}};
    for (auto& c : grids) {
        const Geo g = make_geo(c[0], c[1], c[2]);
        ok = host_check_one<1>(g) && ok;
        ok = host_check_one<2>(g) && ok;
        ok = host_check_one<4>(g) && ok;
    }
    return ok ? 0 : 1;
}

int main(int argc, char** argv) {
    if (argc > 1 && strcmp(argv[1], "--host-check") == 0) return host_check();
    const int reps = 20;
    int nsm = 0;
    CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
    const Geo g = make_geo(129, 129, 129);
    Dict DP, DR;
    build_P(g, DP); build_R(g, DR);
    const long long pad = ((long long)g.n1 * g.n2 + g.n1 + 2 + 1) & ~1LL;
    std::vector<double> hf(g.nf), hc(g.NC);
    srand(1);
    for (auto& v : hf) v = rand() / (double)RAND_MAX;
    for (auto& v : hc) v = rand() / (double)RAND_MAX;
    auto up_dict = [&](const Dict& D, uint16_t*& pid, int*& c0, Pat*& pat, Ent*& ent) {
        CK(cudaMalloc(&pid, D.pid.size() * 2)); CK(cudaMalloc(&c0, D.c0.size() * 4));
        CK(cudaMalloc(&pat, D.pat.size() * sizeof(Pat))); CK(cudaMalloc(&ent, D.ent.size() * sizeof(Ent)));
        CK(cudaMemcpy(pid, D.pid.data(), D.pid.size() * 2, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(c0, D.c0.data(), D.c0.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(pat, D.pat.data(), D.pat.size() * sizeof(Pat), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(ent, D.ent.data(), D.ent.size() * sizeof(Ent), cudaMemcpyHostToDevice));
    };
    uint16_t *pidP, *pidR; int *c0P, *c0R; Pat *patP, *patR; Ent *entP, *entR;
    up_dict(DP, pidP, c0P, patP, entP); up_dict(DR, pidR, c0R, patR, entR);
    double *dfb, *dcb, *dyc, *dyf, *dyf0;
    CK(cudaMalloc(&dfb, (g.nf + 2 * pad) * 8)); CK(cudaMalloc(&dcb, (g.NC + 2 * pad) * 8));
    CK(cudaMalloc(&dyc, g.NC * 8)); CK(cudaMalloc(&dyf, g.nf * 8)); CK(cudaMalloc(&dyf0, g.nf * 8));
    CK(cudaMemset(dfb, 0, (g.nf + 2 * pad) * 8)); CK(cudaMemset(dcb, 0, (g.NC + 2 * pad) * 8));
    double *df = dfb + pad, *dc = dcb + pad;
    CK(cudaMemcpy(df, hf.data(), g.nf * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dc, hc.data(), g.NC * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dyf0, hf.data(), g.nf * 8, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    // restriction: format bytes = fine r (8 B per fine row) + pid, c0 (6 B) and the result (8 B) per coarse row
    {
        std::vector<double> href(g.NC), hy(g.NC);
        const char* names[] = {"reference 1 row/thread", "lines R=1", "lines R=2", "lines R=4"};
        for (int v = 0; v < 4; ++v) {
            auto launch = [&]() {
                switch (v) {
                    case 0: k_ref<<<(int)((g.NC + 255) / 256), 256>>>(g.NC, 0, pidR, c0R, patR, entR, df, dyc); break;
                    case 1: k_restrict_lines<1><<<nsm * 8, 256>>>(g, pidR, c0R, patR, entR, df, dyc); break;
                    case 2: k_restrict_lines<2><<<nsm * 6, 256>>>(g, pidR, c0R, patR, entR, df, dyc); break;
                    case 3: k_restrict_lines<4><<<nsm * 3, 256>>>(g, pidR, c0R, patR, entR, df, dyc); break;
                }
            };
            CK(cudaMemset(dyc, 0, g.NC * 8));
            for (int w = 0; w < 3; ++w) launch();
            CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0));
            for (int r = 0; r < reps; ++r) launch();
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            CK(cudaMemcpy(hy.data(), dyc, g.NC * 8, cudaMemcpyDeviceToHost));
            bool same = true;
            if (v == 0) href = hy; else same = memcmp(href.data(), hy.data(), g.NC * 8) == 0;
            const double us = 1e3 * ms / reps, bytes = 8.0 * g.nf + 14.0 * g.NC;
            printf("restriction 257^3 -> 129^3  %-24s %8.1f us  %7.1f GB/s  %s\n", names[v], us, bytes / us / 1e3, same ? "bit-identical" : "MISMATCH");
        }
    }
    // prolongation x_f += P x_c: every timed launch adds once more, so compare one launch from the same start vector
    {
        std::vector<double> href(g.nf), hy(g.nf);
        const char* names[] = {"reference 1 row/thread", "lines R=1", "lines R=2", "lines R=4"};
        for (int v = 0; v < 4; ++v) {
            auto launch = [&]() {
                switch (v) {
                    case 0: k_ref<<<(int)((g.nf + 255) / 256), 256>>>(g.nf, 1, pidP, c0P, patP, entP, dc, dyf); break;
                    case 1: k_prolong_lines<1><<<nsm * 8, 256>>>(g, pidP, c0P, patP, entP, dc, dyf); break;
                    case 2: k_prolong_lines<2><<<nsm * 6, 256>>>(g, pidP, c0P, patP, entP, dc, dyf); break;
                    case 3: k_prolong_lines<4><<<nsm * 3, 256>>>(g, pidP, c0P, patP, entP, dc, dyf); break;
                }
            };
            CK(cudaMemcpy(dyf, dyf0, g.nf * 8, cudaMemcpyDeviceToDevice));
            launch();
            CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(hy.data(), dyf, g.nf * 8, cudaMemcpyDeviceToHost));
            bool same = true;
            if (v == 0) href = hy; else same = memcmp(href.data(), hy.data(), g.nf * 8) == 0;
            for (int w = 0; w < 2; ++w) launch();
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0));
            for (int r = 0; r < reps; ++r) launch();
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            const double us = 1e3 * ms / reps, bytes = 22.0 * g.nf + 8.0 * g.NC;
            printf("prolongation 129^3 -> 257^3 %-24s %8.1f us  %7.1f GB/s  %s\n", names[v], us, bytes / us / 1e3, same ? "bit-identical" : "MISMATCH");
        }
    }
    return 0;
}
