// Shared definitions for the mgb200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

namespace mgb200 {

// ---------------------------------------------------------------------------------------------
// errors: every CUDA failure becomes a C++ exception; the C ABI turns it into a status code.
// ---------------------------------------------------------------------------------------------
struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define MGB_CUDA(call)                                                                          \
    do {                                                                                        \
        cudaError_t e_ = (call);                                                                \
        if (e_ != cudaSuccess)                                                                  \
            throw ::mgb200::Error(-2, std::string("CUDA error: ") + cudaGetErrorString(e_) +    \
                                          " at " + __FILE__ + ":" + std::to_string(__LINE__));  \
    } while (0)

#define MGB_CHECK(cond, msg)                                                                    \
    do {                                                                                        \
        if (!(cond)) throw ::mgb200::Error(-1, std::string("mgb200: ") + (msg));                \
    } while (0)

// ---------------------------------------------------------------------------------------------
// value types.  ComplexF64 is a plain pair of doubles with the textbook product
// (ar*br - ai*bi, ar*bi + ai*br): the same operation order the CPU oracle (C99 complex,
// -ffp-contract=off) uses, so thread-per-row results are reproducible bit for bit.
// The library is compiled with -fmad=false for the same reason.
// ---------------------------------------------------------------------------------------------
struct __align__(16) cplx {
    double x, y;
};

__host__ __device__ __forceinline__ cplx make_cplx(double a, double b) {
    cplx c;
    c.x = a;
    c.y = b;
    return c;
}
__host__ __device__ __forceinline__ cplx operator+(cplx a, cplx b) { return make_cplx(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ cplx operator-(cplx a, cplx b) { return make_cplx(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ cplx operator-(cplx a) { return make_cplx(-a.x, -a.y); }
__host__ __device__ __forceinline__ cplx operator*(cplx a, cplx b) {
    return make_cplx(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__host__ __device__ __forceinline__ cplx operator*(double a, cplx b) { return make_cplx(a * b.x, a * b.y); }
__host__ __device__ __forceinline__ cplx operator*(cplx b, double a) { return make_cplx(a * b.x, a * b.y); }
__host__ __device__ __forceinline__ cplx conj_(cplx a) { return make_cplx(a.x, -a.y); }
__host__ __device__ __forceinline__ double conj_(double a) { return a; }
__host__ __device__ __forceinline__ double abs2(cplx a) { return a.x * a.x + a.y * a.y; }
__host__ __device__ __forceinline__ double abs2(double a) { return a * a; }

// complex division (Smith's algorithm is not needed for the Krylov scalars; plain formula)
__host__ __device__ __forceinline__ cplx operator/(cplx a, cplx b) {
    double den = b.x * b.x + b.y * b.y;
    return make_cplx((a.x * b.x + a.y * b.y) / den, (a.y * b.x - a.x * b.y) / den);
}


// Single-precision twins (Float32 / ComplexF32 hierarchies of the reference, `singlePrecision` in MGdef.jl:119,151:
// the cycle then runs in single precision under a double-precision Krylov method, SolveFuncs.jl:52-60).
struct __align__(8) cplxf {
    float x, y;
};
__host__ __device__ __forceinline__ cplxf make_cplxf(float a, float b) {
    cplxf c;
    c.x = a;
    c.y = b;
    return c;
}
__host__ __device__ __forceinline__ cplxf operator+(cplxf a, cplxf b) { return make_cplxf(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ cplxf operator-(cplxf a, cplxf b) { return make_cplxf(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ cplxf operator-(cplxf a) { return make_cplxf(-a.x, -a.y); }
__host__ __device__ __forceinline__ cplxf operator*(cplxf a, cplxf b) {
    return make_cplxf(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__host__ __device__ __forceinline__ cplxf operator*(float a, cplxf b) { return make_cplxf(a * b.x, a * b.y); }
__host__ __device__ __forceinline__ cplxf operator*(cplxf b, float a) { return make_cplxf(a * b.x, a * b.y); }
__host__ __device__ __forceinline__ cplxf operator*(double a, cplxf b) { return make_cplxf((float)a * b.x, (float)a * b.y); }
__host__ __device__ __forceinline__ cplxf operator*(cplxf b, double a) { return make_cplxf((float)a * b.x, (float)a * b.y); }
__host__ __device__ __forceinline__ cplxf conj_(cplxf a) { return make_cplxf(a.x, -a.y); }
__host__ __device__ __forceinline__ float conj_(float a) { return a; }
__host__ __device__ __forceinline__ double abs2(cplxf a) { return (double)a.x * a.x + (double)a.y * a.y; }
__host__ __device__ __forceinline__ double abs2(float a) { return (double)a * a; }
__host__ __device__ __forceinline__ cplxf operator/(cplxf a, cplxf b) {
    float den = b.x * b.x + b.y * b.y;
    return make_cplxf((a.x * b.x + a.y * b.y) / den, (a.y * b.x - a.x * b.y) / den);
}

template <typename T>
struct VT;
template <>
struct VT<double> {
    typedef double real_t;
    static constexpr bool is_complex = false;
    __host__ __device__ static __forceinline__ double zero() { return 0.0; }
    __host__ __device__ static __forceinline__ double one() { return 1.0; }
    __host__ __device__ static __forceinline__ double from_real(double a) { return a; }
    __host__ __device__ static __forceinline__ double re(double a) { return a; }
    __host__ __device__ static __forceinline__ double im(double) { return 0.0; }
    __host__ __device__ static __forceinline__ double make(double a, double) { return a; }
};
template <>
struct VT<cplx> {
    typedef double real_t;
    static constexpr bool is_complex = true;
    __host__ __device__ static __forceinline__ cplx zero() { return make_cplx(0.0, 0.0); }
    __host__ __device__ static __forceinline__ cplx one() { return make_cplx(1.0, 0.0); }
    __host__ __device__ static __forceinline__ cplx from_real(double a) { return make_cplx(a, 0.0); }
    __host__ __device__ static __forceinline__ double re(cplx a) { return a.x; }
    __host__ __device__ static __forceinline__ double im(cplx a) { return a.y; }
    __host__ __device__ static __forceinline__ cplx make(double a, double b) { return make_cplx(a, b); }
};

template <>
struct VT<float> {
    typedef float real_t;
    static constexpr bool is_complex = false;
    __host__ __device__ static __forceinline__ float zero() { return 0.0f; }
    __host__ __device__ static __forceinline__ float one() { return 1.0f; }
    __host__ __device__ static __forceinline__ float from_real(double a) { return (float)a; }
    __host__ __device__ static __forceinline__ double re(float a) { return a; }
    __host__ __device__ static __forceinline__ double im(float) { return 0.0; }
    __host__ __device__ static __forceinline__ float make(double a, double) { return (float)a; }
};
template <>
struct VT<cplxf> {
    typedef float real_t;
    static constexpr bool is_complex = true;
    __host__ __device__ static __forceinline__ cplxf zero() { return make_cplxf(0.0f, 0.0f); }
    __host__ __device__ static __forceinline__ cplxf one() { return make_cplxf(1.0f, 0.0f); }
    __host__ __device__ static __forceinline__ cplxf from_real(double a) { return make_cplxf((float)a, 0.0f); }
    __host__ __device__ static __forceinline__ double re(cplxf a) { return a.x; }
    __host__ __device__ static __forceinline__ double im(cplxf a) { return a.y; }
    __host__ __device__ static __forceinline__ cplxf make(double a, double b) { return make_cplxf((float)a, (float)b); }
};

// Wide<T>: the double-precision twin of a value type.  The coarsest-grid factorisation of a single-precision hierarchy
// is held in double precision, as Julia's lu() does for Float32 / ComplexF32 sparse matrices (UMFPACK has no single
// precision: the matrix is promoted, `LU \ b` returns double precision and `x[:] = z` rounds, MGcycle.jl:176-179).
template <typename T> struct Wide { typedef T type; };
template <> struct Wide<float> { typedef double type; };
template <> struct Wide<cplxf> { typedef cplx type; };
__host__ __device__ __forceinline__ double widen(double a) { return a; }
__host__ __device__ __forceinline__ cplx widen(cplx a) { return a; }
__host__ __device__ __forceinline__ double widen(float a) { return (double)a; }
__host__ __device__ __forceinline__ cplx widen(cplxf a) { return make_cplx((double)a.x, (double)a.y); }
__host__ __device__ __forceinline__ void narrow(double a, double& o) { o = a; }
__host__ __device__ __forceinline__ void narrow(cplx a, cplx& o) { o = a; }
__host__ __device__ __forceinline__ void narrow(double a, float& o) { o = (float)a; }
__host__ __device__ __forceinline__ void narrow(cplx a, cplxf& o) { o = make_cplxf((float)a.x, (float)a.y); }

// read-only (non-coherent) loads
__device__ __forceinline__ double ldg_(const double* p) { return __ldg(p); }
__device__ __forceinline__ cplx ldg_(const cplx* p) {
    double2 v = __ldg(reinterpret_cast<const double2*>(p));
    return make_cplx(v.x, v.y);
}
__device__ __forceinline__ int ldg_(const int* p) { return __ldg(p); }
__device__ __forceinline__ float ldg_(const float* p) { return __ldg(p); }
__device__ __forceinline__ cplxf ldg_(const cplxf* p) {
    float2 v = __ldg(reinterpret_cast<const float2*>(p));
    return make_cplxf(v.x, v.y);
}

__device__ __forceinline__ double shfl_xor_(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
__device__ __forceinline__ cplx shfl_xor_(cplx v, int m) {
    return make_cplx(__shfl_xor_sync(0xffffffffu, v.x, m), __shfl_xor_sync(0xffffffffu, v.y, m));
}
__device__ __forceinline__ double shfl_down_(double v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }
__device__ __forceinline__ cplx shfl_down_(cplx v, int d) {
    return make_cplx(__shfl_down_sync(0xffffffffu, v.x, d), __shfl_down_sync(0xffffffffu, v.y, d));
}

__device__ __forceinline__ float shfl_xor_(float v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
__device__ __forceinline__ cplxf shfl_xor_(cplxf v, int m) {
    return make_cplxf(__shfl_xor_sync(0xffffffffu, v.x, m), __shfl_xor_sync(0xffffffffu, v.y, m));
}
__device__ __forceinline__ float shfl_down_(float v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }
__device__ __forceinline__ cplxf shfl_down_(cplxf v, int d) {
    return make_cplxf(__shfl_down_sync(0xffffffffu, v.x, d), __shfl_down_sync(0xffffffffu, v.y, d));
}

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------------------------------------
// TMA (bulk async copy) + mbarrier helpers: global -> shared staging of CSR chunks.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    const uint32_t a = smem_u32(bar);
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(a), "r"(parity)
            : "memory");
    } while (!ok);
}

}  // namespace mgb200
