// Halo exchange and coarse gather over NVLink peer memory (one process per GPU, CUDA IPC).
//
// every operator application on a row-partitioned level needs one ghost plane of
// its input vector from each slab neighbour.  Instead of ncclSend/ncclRecv pairs (a pack kernel
// plus a NCCL group: ~25 us per exchange, more than a coarse-level kernel) the ranks write
// straight into each other's memory:
//
//   put   : every rank packs the rows its neighbours need and stores them THROUGH NVLINK into the
//           neighbour's receive buffer (mapped with cudaIpcOpenMemHandle); the last CTA to finish
//           publishes the exchange number in the neighbour's flag word (fence.sys + st.release.sys).
//   wait  : the CTAs then poll the local flag words until every source has published this exchange
//           and copy the receive buffer into the ghost rows of the vector.
//
// Put and wait are one ordinary kernel on the hierarchy's stream, so whole V/F/W cycles - exchanges included -
// are captured into one CUDA graph per rank.  The exchange number lives in device memory and is
// advanced by the put kernel itself (a replayed graph cannot carry it as a launch argument).
// Receive buffers are double-buffered on the parity of the exchange number: a rank can only start
// put e+2 after its wait e+1, i.e. after the neighbour's put e+1, which follows the neighbour's
// wait e in stream order - so buffer (e mod 2) is free again.
//
// One channel per distributed level (halo of that level's vectors) and one for the gather of the
// first replicated level (every rank restricts its own coarse rows and broadcasts the piece).
// If the peers are not IPC/P2P reachable the hierarchy keeps the NCCL path (solver.cuh).
#pragma once
#include "dist.cuh"

namespace mgb200 {

constexpr int P2P_MAXW = 16;
constexpr int P2P_TRACE_ROWS = 1024;   // MGB200_P2P_TRACE=1: per exchange {start, published, peers arrived, end} of CTA 0

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// device-resident description of one channel
template <typename TV>
struct ChanDev {
    TV* dst[2][P2P_MAXW];                    // where MY rows land in peer q's receive buffer (parity 0/1)
    unsigned long long* flag_dst[P2P_MAXW];  // flag word in peer q's block that I publish to
    const unsigned long long* flag_src[P2P_MAXW];  // local flag word written by source p
    const TV* rbuf[2];                       // my receive buffer (parity 0/1)
    int send_off[P2P_MAXW + 1];              // halo: rows sent to peer q = send_idx[send_off[q] .. send_off[q+1])
    int recv_cnt[P2P_MAXW];                  // halo: ghosts received from p; gather: rows of p's piece
    int recv_off[P2P_MAXW];                  // halo: first ghost of p; gather: first row of p's piece
    int world, rank;
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// receive buffers are written by the peer: read them around L1
__device__ __forceinline__ double ld_cv(const double* p) { return __ldcv(p); }
__device__ __forceinline__ cplx ld_cv(const cplx* p) {
    const double2 v = __ldcv(reinterpret_cast<const double2*>(p));
    return make_cplx(v.x, v.y);
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Publish exchange number e to every destination once ALL CTAs of the kernel have stored their rows: every CTA
// orders its stores with one system-scope fence behind a CTA barrier and takes a ticket; the last one publishes.
template <typename TV>
__device__ __forceinline__ void p2p_publish(const ChanDev<TV>* cd, unsigned long long e, unsigned long long* epoch,
                                            unsigned* ticket, bool bcast) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        if (atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1) {
            __threadfence_system();
            for (int q = 0; q < cd->world; ++q) {
                if (q == cd->rank) continue;
                if (bcast || cd->send_off[q + 1] > cd->send_off[q]) st_release_sys(cd->flag_dst[q], e);
            }
            *epoch = e;
        }
    }
}
// Wait until every source has published exchange e (one polling thread per CTA, then a CTA barrier).
template <typename TV>
__device__ __forceinline__ void p2p_wait(const ChanDev<TV>* cd, unsigned long long e) {
    if (threadIdx.x == 0) {
        for (int p = 0; p < cd->world; ++p)
            if (p != cd->rank && cd->recv_cnt[p] > 0)
                while (ld_acquire_sys(cd->flag_src[p]) < e) {
                }
    }
    __syncthreads();
}

// Halo exchange in one kernel: put (gather the owned rows the peers asked for and store them into the peers'
// buffers), publish, wait for the peers' rows, unpack ghost g of the receive buffer to row ghost_pos(g).
// The grid is small (<= 64 CTAs), so all CTAs are resident and the polling CTAs cannot starve the publisher.
template <typename TV>
__global__ void p2p_halo_kernel(const ChanDev<TV>* __restrict__ cd, TV* __restrict__ v,
                                const int* __restrict__ send_idx, int n_send, long long n_ghost, long long n_lo,
                                long long n_owned, int m, unsigned long long* epoch, unsigned* ticket,
                                unsigned long long* trace) {
    const unsigned long long e = *epoch + 1;
    const int par = (int)(e & 1);
    const bool tr = trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
    unsigned long long* trow = trace + (e % P2P_TRACE_ROWS) * 4;
    if (tr) trow[0] = globaltimer_ns();
    const long long total = (long long)n_send * m;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t / m), j = (int)(t % m);
        int q = 0;
        while (i >= cd->send_off[q + 1]) ++q;
        cd->dst[par][q][(long long)(i - cd->send_off[q]) * m + j] = v[(long long)send_idx[i] * m + j];
    }
    p2p_publish(cd, e, epoch, ticket, false);
    if (tr) trow[1] = globaltimer_ns();
    p2p_wait(cd, e);
    if (tr) trow[2] = globaltimer_ns();
    const TV* rb = cd->rbuf[par];
    const long long tot2 = n_ghost * m;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < tot2;
         t += (long long)gridDim.x * blockDim.x) {
        const long long g = t / m;
        const int j = (int)(t % m);
        const long long pos = g < n_lo ? g - n_lo : n_owned + (g - n_lo);
        v[pos * m + j] = ld_cv(rb + t);
    }
    if (tr) trow[3] = globaltimer_ns();
}

// Gather of a replicated vector in one kernel: my piece v[off .. off+cnt) (element units) goes to the same place
// of every peer's buffer; after the wait the pieces of the other ranks are copied out of my buffer.
template <typename TV>
__global__ void p2p_gather_kernel(const ChanDev<TV>* __restrict__ cd, TV* __restrict__ v, long long off,
                                  long long cnt, int m, unsigned long long* epoch, unsigned* ticket) {
    const unsigned long long e = *epoch + 1;
    const int par = (int)(e & 1);
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < cnt;
         t += (long long)gridDim.x * blockDim.x) {
        const TV val = v[off + t];
        for (int q = 0; q < cd->world; ++q)
            if (q != cd->rank) cd->dst[par][q][t] = val;
    }
    p2p_publish(cd, e, epoch, ticket, true);
    p2p_wait(cd, e);
    const TV* rb = cd->rbuf[par];
    for (int p = 0; p < cd->world; ++p) {
        if (p == cd->rank) continue;
        const long long o = (long long)cd->recv_off[p] * m, c = (long long)cd->recv_cnt[p] * m;
        for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < c;
             t += (long long)gridDim.x * blockDim.x)
            v[o + t] = ld_cv(rb + o + t);
    }
}

// host side of one channel
struct ChanHost {
    bool used = false;
    long long rows = 0;        // rows of the receive buffer (per parity): n_ghost (halo) or nc_global (gather)
    size_t buf_off[2] = {0, 0};  // byte offsets of the receive buffers inside my block
    void* dev = nullptr;       // ChanDev<TV> on the device
};

struct P2P {
    bool on = false;
    unsigned char* block = nullptr;          // my IPC-exported block: flag words, then the receive buffers
    size_t block_bytes = 0;
    std::vector<unsigned char*> peer;        // mapped blocks of the peers (peer[rank] == block)
    std::vector<ChanHost> chan;              // levels halo channels + 1 gather channel
    unsigned long long* epoch = nullptr;     // per channel, device
    unsigned* ticket = nullptr;              // per channel, device
    int gather_level = -1;                   // level index (0-based) of the first replicated level
    unsigned long long* trace = nullptr;     // per halo channel P2P_TRACE_ROWS x 4 timestamps (MGB200_P2P_TRACE=1)
};

}  // namespace mgb200
