// Halo exchange and coarse gather over NVLink peer memory (one process per GPU, CUDA IPC).
//
// every operator application on a row-partitioned level needs one ghost plane of
// its input vector from each slab neighbour.  Instead of ncclSend/ncclRecv pairs (a pack kernel
// plus a NCCL group: ~25 us per exchange, more than a coarse-level kernel) the ranks write
// straight into each other's memory:
//
//   put   : every rank packs the rows its neighbours need and stores them THROUGH NVLINK into the
//           neighbour's receive buffer (mapped with cudaIpcOpenMemHandle) as self-validating 8-byte words:
//           4 bytes of payload + the exchange number.
//   get   : the same kernel then polls its own receive buffer word by word (the exchange number in the word says
//           the payload has arrived) and writes the ghost rows of the vector.
//
// No fence, no flag, no ordering between words is needed - an aligned 8-byte store is single-copy atomic - so an
// exchange costs one NVLink store latency instead of store + fence round trip + flag.  (First version of this
// file: data, fence.sys, flag with st.release.sys, poll with ld.acquire.sys: 15-30 us per exchange at N = 4-8,
// profiles/r01c_p2p_trace_n4.log.)
//
// Put and get are one ordinary kernel on the hierarchy's stream, so whole V/F/W cycles - exchanges included -
// are captured into one CUDA graph per rank.  The exchange number lives in device memory and is
// advanced by the kernel itself (a replayed graph cannot carry it as a launch argument).
// Receive buffers are double-buffered on the parity of the exchange number: a rank can only start
// exchange e+2 after its exchange e+1 completed, i.e. after the neighbour's put e+1, which follows the
// neighbour's get e in stream order - so buffer (e mod 2) is free again, and the words it still holds carry
// exchange number e, never e+2.
//
// One channel per distributed level (halo of that level's vectors) and one for the gather of the
// first replicated level (every rank restricts its own coarse rows and broadcasts the piece).
// If the peers are not IPC/P2P reachable the hierarchy keeps the NCCL path (solver.cuh).
#pragma once
#include "dist.cuh"
#include "ll.cuh"

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

// The exchange number is advanced by the last CTA to finish (every CTA has read it by then).  `consumed` (may be
// null) counts the exchanges whose consumer has passed: a consumer that runs BESIDE the exchange kernel and waits for
// the ghost rows inside the kernel (box.cuh) advances it itself; in the serial form the exchange kernel does.
__device__ __forceinline__ void p2p_advance(unsigned long long e, unsigned long long* epoch, unsigned* ticket,
                                            unsigned long long* consumed = nullptr) {
    __threadfence();          // the ghost rows written above are visible before the exchange number is
    __syncthreads();
    if (threadIdx.x == 0 && atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1) {
        if (consumed) *consumed = e;
        __threadfence();
        *reinterpret_cast<volatile unsigned long long*>(epoch) = e;
    }
}

// Halo exchange in one kernel: gather the owned rows the peers asked for and store them as LL words straight into
// the peers' receive buffers (through NVLink); then poll the own receive buffer word by word and write ghost g to
// row ghost_pos(g).  The grid is small (<= 64 CTAs), all CTAs are resident, nobody waits on another CTA.
template <typename TV>
__global__ void p2p_halo_kernel(const ChanDev<TV>* __restrict__ cd, TV* __restrict__ v,
                                const int* __restrict__ send_idx, int n_send, long long n_ghost, long long n_lo,
                                long long n_owned, int m, unsigned long long* epoch, unsigned* ticket,
                                unsigned long long* trace, int skip_put, unsigned long long* consumed) {
    constexpr int W = LL<TV>::W;
    const unsigned long long e = *epoch + 1;
    const int par = (int)(e & 1);
    const unsigned flag = (unsigned)e;
    const bool tr = trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
    unsigned long long* trow = trace + (e % P2P_TRACE_ROWS) * 4;
    if (tr) trow[0] = globaltimer_ns();
    const long long total = skip_put ? 0 : (long long)n_send * m;   // skip_put: the producing kernel stored the rows
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t / m), j = (int)(t % m);
        int q = 0;
        while (i >= cd->send_off[q + 1]) ++q;
        unsigned long long* dst = reinterpret_cast<unsigned long long*>(cd->dst[par][q]);
        LL<TV>::put(dst + ((long long)(i - cd->send_off[q]) * m + j) * W, v[(long long)send_idx[i] * m + j], flag);
    }
    if (tr) trow[1] = globaltimer_ns();
    const unsigned long long* rb = reinterpret_cast<const unsigned long long*>(cd->rbuf[par]);
    const long long tot2 = n_ghost * m;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < tot2;
         t += (long long)gridDim.x * blockDim.x) {
        const long long g = t / m;
        const int j = (int)(t % m);
        const long long pos = g < n_lo ? g - n_lo : n_owned + (g - n_lo);
        v[pos * m + j] = LL<TV>::get(rb + t * W, flag);
    }
    if (tr) trow[2] = trow[3] = globaltimer_ns();
    p2p_advance(e, epoch, ticket, consumed);
}

// the exchange that ran beside a consumer which, after all, did not wait for it in the kernel
static __global__ void p2p_mark_consumed_kernel(const unsigned long long* epoch, unsigned long long* consumed) { *consumed = *epoch; }

// Gather of a replicated vector in one kernel: my piece v[off .. off+cnt) (element units) goes to the same place
// of every peer's buffer; the pieces of the other ranks are polled out of my buffer (same layout as v).
template <typename TV>
__global__ void p2p_gather_kernel(const ChanDev<TV>* __restrict__ cd, TV* __restrict__ v, long long off,
                                  long long cnt, int m, unsigned long long* epoch, unsigned* ticket) {
    constexpr int W = LL<TV>::W;
    const unsigned long long e = *epoch + 1;
    const int par = (int)(e & 1);
    const unsigned flag = (unsigned)e;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < cnt;
         t += (long long)gridDim.x * blockDim.x) {
        const TV val = v[off + t];
        for (int q = 0; q < cd->world; ++q)
            if (q != cd->rank) LL<TV>::put(reinterpret_cast<unsigned long long*>(cd->dst[par][q]) + t * W, val, flag);
    }
    const unsigned long long* rb = reinterpret_cast<const unsigned long long*>(cd->rbuf[par]);
    for (int p = 0; p < cd->world; ++p) {
        if (p == cd->rank) continue;
        const long long o = (long long)cd->recv_off[p] * m, c = (long long)cd->recv_cnt[p] * m;
        for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < c;
             t += (long long)gridDim.x * blockDim.x)
            v[o + t] = LL<TV>::get(rb + (o + t) * W, flag);
    }
    p2p_advance(e, epoch, ticket);
}

// host side of one channel
struct ChanHost {
    bool used = false;
    PutPlan put = no_put();    // halo channels whose send sets are the two end ranges of the owned rows (fused put)
    long long rows = 0;        // rows of the receive buffer (per parity): n_ghost (halo) or nc_global (gather)
    size_t buf_off[2] = {0, 0};  // byte offsets of the receive buffers inside my block
    void* dev = nullptr;       // ChanDev<TV> on the device
};

struct P2P {
    bool on = false;
    unsigned char* block = nullptr;          // my IPC-exported block: flag words, then the receive buffers
    size_t block_bytes = 0;
    std::vector<unsigned char*> peer;        // mapped blocks of the peers (peer[rank] == block)
    std::vector<char> same_process;          // peer q is a device of this process (plain peer access, no IPC mapping)
    std::vector<ChanHost> chan;              // levels halo channels + 1 gather channel
    unsigned long long* epoch = nullptr;     // per channel, device
    unsigned long long* consumed = nullptr;  // per channel, device: exchanges whose consumer has passed (see p2p_advance)
    unsigned* ticket2 = nullptr;             // per channel, device: CTA counter of a consumer that waits in-kernel
    unsigned* ticket = nullptr;              // per channel, device
    int gather_level = -1;                   // level index (0-based) of the first replicated level
    unsigned long long* trace = nullptr;     // per halo channel P2P_TRACE_ROWS x 4 timestamps (MGB200_P2P_TRACE=1)
};

}  // namespace mgb200
