// A C++ caller of libmgb200.so that uses nothing but include/mgb200.h - the role the reference's Julia `ccall`s play
// (src/Multigrid/parRelax.jl:61-79: raw colptr / rowval / nzval of a SparseMatrixCSC, 1-based Int64 indices).
//
//   abi_client probe                      link + load check: version, and the failure mode of mgb200_create
//   abi_client solve <in.bin> <out.bin>   read a hierarchy (arrays exactly as Julia holds them: 1-based), run solveMG and
//                                         solveCG_MG through the C ABI, write the residual histories and solutions
//
// File format of <in.bin> (little endian): int64 levels, cycle ('V'...), maxit; double tol; then per level l < levels:
// int64 n, nc, nnzA, nnzP, nnzR; A colptr[n+1], rowval[nnzA] (int64), nzval[nnzA] (double); P colptr[n+1], rowval, nzval;
// R colptr[nc+1], rowval, nzval; d[n]; then the coarsest matrix: int64 n, nnz; colptr, rowval, nzval; then b[n1].
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "mgb200.h"

static void die(const char* what) {
    std::fprintf(stderr, "abi_client: %s: %s\n", what, mgb200_last_error());
    std::exit(2);
}
template <typename T>
static std::vector<T> rd(FILE* f, size_t n) {
    std::vector<T> v(n);
    if (n && std::fread(v.data(), sizeof(T), n, f) != n) {
        std::fprintf(stderr, "abi_client: short read\n");
        std::exit(3);
    }
    return v;
}
struct Csc {
    std::vector<int64_t> colptr, rowval;
    std::vector<double> nzval;
};
static Csc rd_csc(FILE* f, int64_t ncol, int64_t nnz) {
    Csc m;
    m.colptr = rd<int64_t>(f, ncol + 1);
    m.rowval = rd<int64_t>(f, nnz);
    m.nzval = rd<double>(f, nnz);
    if (m.colptr[0] != 1 || m.colptr[ncol] != nnz + 1) {   // the arrays must be Julia's: 1-based
        std::fprintf(stderr, "abi_client: colptr is not 1-based\n");
        std::exit(3);
    }
    return m;
}

int main(int argc, char** argv) {
    if (argc >= 2 && std::strcmp(argv[1], "probe") == 0) {
        std::printf("version %d\n", mgb200_version());
        mgb200_handle h = nullptr;
        const int64_t pre[2] = {2, 2}, post[2] = {2, 2};
        const int st = mgb200_create(&h, MGB200_FP64, 2, 1, 'V', MGB200_RELAX_DIAG, pre, post, 0);
        std::printf("create status %d%s%s\n", st, st ? ": " : "", st ? mgb200_last_error() : "");
        if (st == 0) mgb200_destroy(h);
        // a null handle is an error, not a crash
        int it = 0;
        double res[4];
        const int st2 = mgb200_solveMG(nullptr, nullptr, nullptr, 1e-6, 3, &it, res);
        std::printf("null-handle status %d\n", st2);
        return st2 == -1 ? 0 : 1;
    }
    if (argc < 4 || std::strcmp(argv[1], "solve") != 0) {
        std::fprintf(stderr, "usage: abi_client probe | solve <in.bin> <out.bin>\n");
        return 1;
    }
    FILE* f = std::fopen(argv[2], "rb");
    if (!f) return 3;
    const std::vector<int64_t> head = rd<int64_t>(f, 3);
    const int levels = (int)head[0];
    const char cycle = (char)head[1];
    const int maxit = (int)head[2];
    const double tol = rd<double>(f, 1)[0];
    std::vector<int64_t> pre(levels, 2), post(levels, 2);
    mgb200_handle h = nullptr;
    if (mgb200_create(&h, MGB200_FP64, levels, 1, cycle, MGB200_RELAX_DIAG, pre.data(), post.data(), 0)) die("create");
    int64_t n1 = 0;
    for (int l = 1; l < levels; ++l) {
        const std::vector<int64_t> sz = rd<int64_t>(f, 5);
        const int64_t n = sz[0], nc = sz[1];
        if (l == 1) n1 = n;
        const Csc A = rd_csc(f, n, sz[2]), P = rd_csc(f, n, sz[3]), R = rd_csc(f, nc, sz[4]);
        const std::vector<double> d = rd<double>(f, n);
        if (mgb200_upload_level(h, l, n, nc, A.colptr.data(), A.rowval.data(), A.nzval.data(), P.colptr.data(), P.rowval.data(),
                                P.nzval.data(), R.colptr.data(), R.rowval.data(), R.nzval.data(), d.data(), /*index_base=*/1))
            die("upload_level");
    }
    {
        const std::vector<int64_t> sz = rd<int64_t>(f, 2);
        const Csc A = rd_csc(f, sz[0], sz[1]);
        if (mgb200_upload_coarsest(h, sz[0], A.colptr.data(), A.rowval.data(), A.nzval.data(), 1)) die("upload_coarsest");
    }
    const std::vector<double> b = rd<double>(f, n1);
    std::fclose(f);

    std::vector<double> x(n1, 0.0), res(maxit + 1, 0.0), xcg(n1, 0.0), rescg(maxit + 1, 0.0);
    int iter = 0, itcg = 0, flag = 0;
    if (mgb200_solveMG(h, b.data(), x.data(), tol, maxit, &iter, res.data())) die("solveMG");
    if (mgb200_solveCG(h, b.data(), xcg.data(), tol, maxit, &itcg, &flag, rescg.data())) die("solveCG");
    if (mgb200_destroy(h)) die("destroy");

    FILE* o = std::fopen(argv[3], "wb");
    if (!o) return 3;
    const int64_t outhead[3] = {iter, itcg, flag};
    std::fwrite(outhead, sizeof(int64_t), 3, o);
    std::fwrite(res.data(), sizeof(double), res.size(), o);
    std::fwrite(rescg.data(), sizeof(double), rescg.size(), o);
    std::fwrite(x.data(), sizeof(double), x.size(), o);
    std::fwrite(xcg.data(), sizeof(double), xcg.size(), o);
    std::fclose(o);
    std::printf("solveMG %d cycles, relres %.3e; solveCG %d iterations, flag %d\n", iter, res[iter] / res[0], itcg, flag);
    return 0;
}
