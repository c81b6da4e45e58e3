// Self-validating 8-byte words for transfers through NVLink peer memory, and the plan a PRODUCING kernel uses to
// store the boundary rows of its result straight into the neighbours' receive buffers (p2p.cuh has the exchange).
#pragma once
#include "common.cuh"

namespace mgb200 {

// ---- LL ("low latency") words: every 8-byte word carries 4 bytes of payload and the exchange number ----------
// An aligned 8-byte store is single-copy atomic, also across NVLink, so the receiver can poll the word itself:
// no fence, no separate flag, no ordering between words.  (The same idea as NCCL's LL protocol.)  A double
// travels as two words, a complex number as four; the receive buffers are twice the size of the data.
__device__ __forceinline__ void st_ll(unsigned long long* p, unsigned lo, unsigned flag) {
    const unsigned long long w = (unsigned long long)lo | ((unsigned long long)flag << 32);
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ unsigned ld_ll(const unsigned long long* p, unsigned flag) {
    unsigned long long w;
    do {
        asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
    } while ((unsigned)(w >> 32) != flag);
    return (unsigned)w;
}
template <typename TV>
struct LL;
template <>
struct LL<double> {
    static constexpr int W = 2;   // words per element
    __device__ static __forceinline__ void put(unsigned long long* dst, double v, unsigned flag) {
        st_ll(dst, (unsigned)__double2loint(v), flag);
        st_ll(dst + 1, (unsigned)__double2hiint(v), flag);
    }
    __device__ static __forceinline__ double get(const unsigned long long* src, unsigned flag) {
        const unsigned lo = ld_ll(src, flag), hi = ld_ll(src + 1, flag);
        return __hiloint2double((int)hi, (int)lo);
    }
};
template <>
struct LL<cplx> {
    static constexpr int W = 4;
    __device__ static __forceinline__ void put(unsigned long long* dst, cplx v, unsigned flag) {
        LL<double>::put(dst, v.x, flag);
        LL<double>::put(dst + 2, v.y, flag);
    }
    __device__ static __forceinline__ cplx get(const unsigned long long* src, unsigned flag) {
        const double a = LL<double>::get(src, flag), b = LL<double>::get(src + 2, flag);
        return make_cplx(a, b);
    }
};

template <>
struct LL<float> {
    static constexpr int W = 1;
    __device__ static __forceinline__ void put(unsigned long long* dst, float v, unsigned flag) {
        st_ll(dst, __float_as_uint(v), flag);
    }
    __device__ static __forceinline__ float get(const unsigned long long* src, unsigned flag) {
        return __uint_as_float(ld_ll(src, flag));
    }
};
template <>
struct LL<cplxf> {
    static constexpr int W = 2;
    __device__ static __forceinline__ void put(unsigned long long* dst, cplxf v, unsigned flag) {
        st_ll(dst, __float_as_uint(v.x), flag);
        st_ll(dst + 1, __float_as_uint(v.y), flag);
    }
    __device__ static __forceinline__ cplxf get(const unsigned long long* src, unsigned flag) {
        const unsigned a = ld_ll(src, flag), b = ld_ll(src + 1, flag);
        return make_cplxf(__uint_as_float(a), __uint_as_float(b));
    }
};

// Fused put: the rows of a level vector that the slab neighbours need (for a z-slab: the first plane goes to the lower
// neighbour, the last plane to the upper one) are stored as LL words by the kernel that COMPUTES them, so the
// transfer rides under the rest of that kernel and the exchange kernel that follows only has to poll and unpack.
// The exchange number comes from device memory (the previous exchange of the channel advanced it), so the plan is a
// constant of the captured graph.  One right-hand side.
struct PutPlan {
    int on;                          // 0: nothing to put
    int lo_cnt;                      // rows [0, lo_cnt) go to the lower neighbour
    int hi_start, hi_cnt;            // rows [hi_start, hi_start + hi_cnt) go to the upper neighbour
    unsigned long long* dst_lo[2];   // landing zones of those rows in the neighbours' receive buffers, per parity
    unsigned long long* dst_hi[2];
    const unsigned long long* epoch; // exchange number of the channel (the put belongs to exchange *epoch + 1)
};
static inline PutPlan no_put() {
    PutPlan p;
    p.on = 0;
    p.lo_cnt = 0;
    p.hi_start = 0x7fffffff;
    p.hi_cnt = 0;
    p.dst_lo[0] = p.dst_lo[1] = p.dst_hi[0] = p.dst_hi[1] = nullptr;
    p.epoch = nullptr;
    return p;
}
template <typename TV>
__device__ __forceinline__ void ll_put_edge(const PutPlan& P, long long row, TV val) {
    if (row < P.lo_cnt || row >= P.hi_start) {
        constexpr int W = LL<TV>::W;
        const unsigned long long e = *P.epoch + 1;
        const int par = (int)(e & 1);
        if (row < P.lo_cnt) LL<TV>::put(P.dst_lo[par] + row * W, val, (unsigned)e);
        if (row >= P.hi_start && row < (long long)P.hi_start + P.hi_cnt)
            LL<TV>::put(P.dst_hi[par] + (row - P.hi_start) * W, val, (unsigned)e);
    }
}

}  // namespace mgb200
