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
// NOT part of the library; not yet run on a B200 (written in a session whose GPU budget was spent; the host check
// passes).
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

// one thread: column i of the R lines [y0, y0+R) of plane z.  x must be readable S2 + S + 1 elements beyond both ends.
template <int R>
__host__ __device__ inline void lines_thread(const Grid& G, int i, int y0, int z, const uint16_t* pid, const Pat* pat,
                                             const Ent* ent, const double* dp, const double* x, const double* b, double* y) {
    const long long row0 = (long long)z * G.S2 + (long long)y0 * G.S + i;
    const int nr = (G.n2 - y0 < R) ? (G.n2 - y0) : R;          // lines left in this plane
    int p0 = pid[row0];
    bool same = (nr == R);
#pragma unroll
    for (int j = 1; j < R; ++j)
        if (j < nr) same = same && (pid[row0 + (long long)j * G.S] == p0);
    if (!same) {                                                  // mixed patterns or a short group: row by row
        for (int j = 0; j < nr; ++j) {
            const long long row = row0 + (long long)j * G.S;
            y[row] = row_reference(G, row, pid, pat, ent, dp, x, b);
        }
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
        const double* xp = x + row0 + (long long)dz * G.S2;
        // which columns / lines of this plane does the pattern touch?
        const bool col_m = (mz & 0x049) != 0, col_0 = (mz & 0x092) != 0, col_p = (mz & 0x124) != 0;   // dx = -1, 0, +1
        const bool lin_m = (mz & 0x007) != 0, lin_p = (mz & 0x1C0) != 0;                               // dy = -1, +1
        double X[3][R + 2];
#pragma unroll
        for (int l = 0; l < R + 2; ++l) {
            const bool need = (l == 0) ? lin_m : (l == R + 1 ? lin_p : true);
            // dy = 0 entries use l = 1..R; dy = -1 uses 0..R-1; dy = +1 uses 2..R+1
            const double* q = xp + (long long)(l - 1) * G.S;
            X[0][l] = (need && col_m) ? q[-1] : 0.0;
            X[1][l] = (need && col_0) ? q[0] : 0.0;
            X[2][l] = (need && col_p) ? q[1] : 0.0;
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
        const long long row = row0 + (long long)j * G.S;
        if (!have_c) xc[j] = x[row];
        y[row] = xc[j] + d * (b[row] - acc[j]);
    }
}

// persistent grid-stride over the flattened (group, column) index; groups ordered plane by plane so that the three
// planes a group reads stay L2 resident
template <int R>
__global__ void __launch_bounds__(256) k_lines(Grid G, const uint16_t* __restrict__ pid, const Pat* __restrict__ pat,
                                               const Ent* __restrict__ ent, const double* __restrict__ dp,
                                               const double* __restrict__ x, const double* __restrict__ b,
                                               double* __restrict__ y) {
    const int gpp = (G.n2 + R - 1) / R;                       // line groups per plane
    const long long total = (long long)G.n3 * gpp * G.S;
    for (long long f = blockIdx.x * (long long)blockDim.x + threadIdx.x; f < total; f += (long long)gridDim.x * blockDim.x) {
        const long long g = f / G.S;
        const int i = (int)(f - g * G.S);
        const int z = (int)(g / gpp), q = (int)(g - (long long)z * gpp);
        lines_thread<R>(G, i, q * R, z, pid, pat, ent, dp, x, b, y);
    }
}

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
    const long long pad = G.S2 + G.S + 1;
    std::vector<double> xbuf(G.n + 2 * pad, 1e300), b(G.n), yref(G.n), y(G.n, -1.0);   // poison in the padding
    double* x = xbuf.data() + pad;
    srand(7);
    for (long long i = 0; i < G.n; ++i) { x[i] = rand() / (double)RAND_MAX; b[i] = rand() / (double)RAND_MAX; }
    for (long long r = 0; r < G.n; ++r) yref[r] = row_reference(G, r, pid.data(), pat.data(), ent.data(), dp.data(), x, b.data());
    const int gpp = (G.n2 + R - 1) / R;
    for (int z = 0; z < G.n3; ++z)
        for (int q = 0; q < gpp; ++q)
            for (int i = 0; i < G.n1; ++i)
                lines_thread<R>(G, i, q * R, z, pid.data(), pat.data(), ent.data(), dp.data(), x, b.data(), y.data());
    const bool ok = memcmp(y.data(), yref.data(), G.n * sizeof(double)) == 0;
    printf("host check %2d-point %dx%dx%d R=%d: %s\n", pts, G.n1, G.n2, G.n3, R, ok ? "bit-identical" : "MISMATCH");
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
        const long long n = G.n, pad = G.S2 + G.S + 1;
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
                               "lines R=4 x6 CTAs/SM", "lines R=8 x2 CTAs/SM", "lines R=8 x4 CTAs/SM"};
        for (int v = 0; v < 8; ++v) {
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
