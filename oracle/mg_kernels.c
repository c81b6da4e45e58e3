/*
 * TEST INFRASTRUCTURE - NOT PRODUCT CODE.
 *
 * CPU restatement of the array kernels on the reference's multigrid hot path
 * (JuliaInv/Multigrid.jl v0.8.0).  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library.
 *
 * PARITY UNPINNED: the reference cannot run here (no julia, no gfortran) and its
 * own tests hold no golden vectors for this path, so
 * these kernels restate the cited lines and the documented semantics of the
 * un-vendored ParSpMatVec 0.1.1 (Manifest.toml:87-91).
 *
 *   spmatmul_*   src/Multigrid/SpMatMul.jl:4-26
 *                y = beta*y + alpha * AT^H x, AT given as 1-based Int64 CSC.  The
 *                fallback `mul!(target, adjoint(AT), x, alpha, beta)` (:9,:23)
 *                defines the maths; ParSpMatVec parallelises over the columns of
 *                AT (= output rows) with OpenMP, accumulates each row sequentially
 *                in stored order, special-cases beta == 0 (y is not read) and
 *                loops over right-hand-side columns outside the row loop.
 *   addvectors_* src/Multigrid/SpMatMul.jl:29-36   target += alpha*x  (BLAS.axpy!)
 *   scaleadd_*   src/Multigrid/MGcycle.jl:129,134  x .+= d .* r  (d broadcast over columns)
 *   nrm2_*, dot_* LinearAlgebra norm / dot (dot conjugates its first argument)
 *
 * Compiled with -ffp-contract=off: a multiply followed by an add, as the
 * reference's compilers emit for x86-64 without -mfma.
 */
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stddef.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef double complex c64;

static inline void set_threads(int64_t nthreads) {
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads((int)nthreads);
#else
    (void)nthreads;
#endif
}

int64_t oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

#define CONJ_f64(a) (a)
#define CONJ_c64(a) conj(a)

#define DEFINE_SPMATMUL(NAME, TA, TX, CONJA)                                                     \
    void NAME(int64_t ncols, const int64_t *colptr, const int64_t *rowval, const TA *nzval,    \
              const TX *x, int64_t ldx, TX *y, int64_t ldy, int64_t nrhs, const TX *palpha,     \
              const TX *pbeta, int64_t nthreads) {                                               \
        const TX alpha = *palpha, beta = *pbeta;                                                 \
        set_threads(nthreads);                                                                   \
        for (int64_t c = 0; c < nrhs; ++c) {                                                     \
            const TX *xc = x + c * ldx;                                                          \
            TX *yc = y + c * ldy;                                                                \
            if (beta == 0) {                                                                     \
                _Pragma("omp parallel for schedule(static)")                                     \
                for (int64_t j = 0; j < ncols; ++j) {                                            \
                    TX t = 0;                                                                    \
                    for (int64_t k = colptr[j] - 1; k < colptr[j + 1] - 1; ++k)                  \
                        t = t + CONJA(nzval[k]) * xc[rowval[k] - 1];                             \
                    yc[j] = alpha * t;                                                           \
                }                                                                                \
            } else {                                                                             \
                _Pragma("omp parallel for schedule(static)")                                     \
                for (int64_t j = 0; j < ncols; ++j) {                                            \
                    TX t = 0;                                                                    \
                    for (int64_t k = colptr[j] - 1; k < colptr[j + 1] - 1; ++k)                  \
                        t = t + CONJA(nzval[k]) * xc[rowval[k] - 1];                             \
                    yc[j] = beta * yc[j] + alpha * t;                                            \
                }                                                                                \
            }                                                                                    \
        }                                                                                        \
    }

DEFINE_SPMATMUL(spmatmul_f64_f64, double, double, CONJ_f64)
DEFINE_SPMATMUL(spmatmul_c64_c64, c64, c64, CONJ_c64)
DEFINE_SPMATMUL(spmatmul_f64_c64, double, c64, CONJ_f64)

#define DEFINE_VEC(SUF, T, ABS2, CONJ)                                                           \
    void addvectors_##SUF(int64_t n, const T *palpha, const T *x, T *y, int64_t nthreads) {      \
        const T alpha = *palpha;                                                                 \
        set_threads(nthreads);                                                                   \
        _Pragma("omp parallel for schedule(static)")                                             \
        for (int64_t i = 0; i < n; ++i) y[i] = y[i] + alpha * x[i];                              \
    }                                                                                            \
    void scaleadd_##SUF(int64_t n, int64_t nrhs, const T *d, const T *r, T *x,                   \
                        int64_t nthreads) {                                                      \
        set_threads(nthreads);                                                                   \
        for (int64_t c = 0; c < nrhs; ++c) {                                                     \
            const T *rc = r + c * n;                                                             \
            T *xc = x + c * n;                                                                   \
            _Pragma("omp parallel for schedule(static)")                                         \
            for (int64_t i = 0; i < n; ++i) xc[i] = xc[i] + d[i] * rc[i];                        \
        }                                                                                        \
    }                                                                                            \
    double nrm2_##SUF(int64_t n, const T *x, int64_t nthreads) {                                 \
        double s = 0.0;                                                                          \
        set_threads(nthreads);                                                                   \
        _Pragma("omp parallel for schedule(static) reduction(+ : s)")                            \
        for (int64_t i = 0; i < n; ++i) s += ABS2(x[i]);                                         \
        return sqrt(s);                                                                          \
    }                                                                                            \
    void dot_##SUF(int64_t n, const T *x, const T *y, T *out, int64_t nthreads) {                \
        double sr = 0.0, si = 0.0;                                                               \
        set_threads(nthreads);                                                                   \
        _Pragma("omp parallel for schedule(static) reduction(+ : sr, si)")                       \
        for (int64_t i = 0; i < n; ++i) {                                                        \
            c64 p = CONJ(x[i]) * y[i];                                                           \
            sr += creal(p);                                                                      \
            si += cimag(p);                                                                      \
        }                                                                                        \
        *out = (T)(sr + si * I);                                                                 \
    }

#define ABS2_f64(a) ((a) * (a))
#define ABS2_c64(a) (creal(a) * creal(a) + cimag(a) * cimag(a))

DEFINE_VEC(f64, double, ABS2_f64, CONJ_f64)
DEFINE_VEC(c64, c64, ABS2_c64, CONJ_c64)
