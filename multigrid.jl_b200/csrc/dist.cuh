// Multi-GPU row partition: ghost maps, halo exchange and scalar all-reduces over NCCL.
//
// One process per GPU.  A distributed level keeps its vectors as [ghosts below | owned rows | ghosts
// above] with the vector pointer at the first owned row (lower ghosts have negative indices); the
// CSR column indices are remapped into that layout at finalisation, so the compute kernels are
// exactly the single-GPU ones and stencil matrices keep their row-relative structure.  Communication per operator application = one halo exchange of
// the input vector (ncclSend/ncclRecv pairs inside one group, peers are the slab neighbours for
// the z-slab layout of src/DomainDecomposition/DDIndices.jl:41-47) and, for Krylov scalars,
// an in-place ncclAllReduce on a few doubles.  Coarse levels are replicated: the restricted
// residual is all-gathered once per cycle and the coarse part of the cycle runs redundantly
// on every GPU with no further communication (agglomeration).
//
// NCCL is loaded with dlopen only when a distributed run is initialised, so single-GPU users
// carry no NCCL dependency.
#pragma once
#include <dlfcn.h>
#include <unistd.h>
#include <nccl.h>

#include <algorithm>
#include <vector>

#include "hierarchy.cuh"

namespace mgb200 {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;

    void load() {
        if (handle) return;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (handle) break;
        }
        if (!handle) throw Error(-3, std::string("cannot load NCCL: ") + dlerror());
#define MGB_SYM(field, sym)                                                        \
    field = reinterpret_cast<decltype(field)>(dlsym(handle, sym));                 \
    if (!field) throw Error(-3, std::string("NCCL symbol missing: ") + sym);
        MGB_SYM(GetUniqueId, "ncclGetUniqueId")
        MGB_SYM(CommInitRank, "ncclCommInitRank")
        MGB_SYM(CommDestroy, "ncclCommDestroy")
        MGB_SYM(AllReduce, "ncclAllReduce")
        MGB_SYM(Broadcast, "ncclBroadcast")
        MGB_SYM(AllGather, "ncclAllGather")
        MGB_SYM(Send, "ncclSend")
        MGB_SYM(Recv, "ncclRecv")
        MGB_SYM(GroupStart, "ncclGroupStart")
        MGB_SYM(GroupEnd, "ncclGroupEnd")
        MGB_SYM(GetErrorString, "ncclGetErrorString")
#undef MGB_SYM
    }
};

static NcclApi& nccl() {
    static NcclApi api;
    return api;
}

#define MGB_NCCL(call)                                                                             \
    do {                                                                                           \
        ncclResult_t r_ = (call);                                                                  \
        if (r_ != ncclSuccess)                                                                     \
            throw ::mgb200::Error(-3, std::string("NCCL error: ") + nccl().GetErrorString(r_) +    \
                                          " at " + __FILE__ + ":" + std::to_string(__LINE__));     \
    } while (0)

struct Comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    bool active() const { return world > 1; }
};

// host-staged local rows of a distributed matrix: CSR with GLOBAL column indices
template <typename TA>
struct HostRows {
    long long n_rows = 0;
    std::vector<int64_t> rowptr;    // n_rows + 1, base 0
    std::vector<int64_t> col;       // global
    std::vector<TA> val;            // already in operator form (conjugated if needed)
    bool present() const { return !rowptr.empty(); }
    void clear() {
        n_rows = 0;
        rowptr.clear(); rowptr.shrink_to_fit();
        col.clear(); col.shrink_to_fit();
        val.clear(); val.shrink_to_fit();
    }
};

// ---------------------------------------------------------------------------------------------
// host-only planning (no GPU, no NCCL): ghost set of a row slab
// ---------------------------------------------------------------------------------------------
// Collects the sorted unique global column ids outside [lo, hi).
static inline void collect_ghosts(const int64_t* col, long long nnz, long long lo, long long hi,
                                  std::vector<long long>& ghosts) {
    for (long long k = 0; k < nnz; ++k)
        if (col[k] < lo || col[k] >= hi) ghosts.push_back(col[k]);
}
static inline void sort_unique(std::vector<long long>& v) {
    std::sort(v.begin(), v.end());
    v.erase(std::unique(v.begin(), v.end()), v.end());
}
// global id -> local index in the [owned | ghost] layout (host planner of the CPU test-suite)
static inline long long to_local(long long c, long long lo, long long hi, const std::vector<long long>& ghosts) {
    if (c >= lo && c < hi) return c - lo;
    auto it = std::lower_bound(ghosts.begin(), ghosts.end(), c);
    return (hi - lo) + (it - ghosts.begin());
}
// global id -> index relative to the first OWNED row in the [ghosts below lo | owned | ghosts above hi]
// layout the device uses: lower ghosts get negative indices.  For a z-slab the ghost planes are the
// contiguous global ranges next to [lo,hi), so the map is the pure shift c - lo and a stencil matrix keeps
// its row-relative column offsets (the stencil dictionary of pattern.cuh survives the partition).
static inline long long to_local_split(long long c, long long lo, long long hi, const std::vector<long long>& ghosts,
                                       long long n_lo) {
    if (c >= lo && c < hi) return c - lo;
    const long long g = std::lower_bound(ghosts.begin(), ghosts.end(), c) - ghosts.begin();
    return g < n_lo ? g - n_lo : (hi - lo) + (g - n_lo);
}

// Rows [lo, hi) of a staged matrix (columns already in the split layout: < 0 lower ghost, >= n_in_owned upper
// ghost) that read owned rows only.  For a z-slab these are all rows but the first and the last plane.
template <typename TA>
static void interior_rows(const HostRows<TA>& H, long long n_in_owned, int& lo, int& hi) {
    long long lo_max = -1, hi_min = H.n_rows;
    for (long long i = 0; i < H.n_rows; ++i)
        for (long long k = H.rowptr[i]; k < H.rowptr[i + 1]; ++k) {
            if (H.col[k] < 0) lo_max = std::max(lo_max, i);
            else if (H.col[k] >= n_in_owned) hi_min = std::min(hi_min, i);
        }
    lo = (int)(lo_max + 1);
    hi = (int)hi_min;
    if (hi < lo) lo = hi = 0;
}

// vector space of one distributed level
struct DistSpace {
    bool dist = false;
    long long n_global = 0;
    std::vector<long long> row_offsets;  // world + 1
    long long lo = 0, hi = 0;            // owned global range
    long long n_owned = 0, n_ghost = 0;
    long long n_lo = 0;                  // ghosts with global id < lo: they sit IN FRONT of the owned rows
    std::vector<long long> ghosts;       // sorted global ids
    std::vector<int> recv_cnt, recv_off; // per peer (ghosts are grouped by owner because owners hold ranges)
    std::vector<int> send_cnt, send_off;
    int* d_send_idx = nullptr;           // local owned indices to pack, concatenated per peer
    std::vector<int> send_idx_host;      // the same on the host (fused-put planning)
    int n_send = 0;
    void* sendbuf = nullptr;             // n_send * m * sizeof(TV)
    size_t sendbuf_bytes = 0;
    // position of ghost number g relative to the first owned row
    long long ghost_pos(long long g) const { return g < n_lo ? g - n_lo : n_owned + (g - n_lo); }
    int owner_of(long long gid) const {
        auto it = std::upper_bound(row_offsets.begin(), row_offsets.end(), gid);
        return (int)(it - row_offsets.begin()) - 1;
    }
    void release() {
        dev_free(d_send_idx);
        if (sendbuf) cudaFree(sendbuf);
        sendbuf = nullptr;
        sendbuf_bytes = 0;
    }
};

template <typename TV>
__global__ void pack_kernel(const TV* __restrict__ v, const int* __restrict__ idx, int n, int m, TV* __restrict__ out) {
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long total = (long long)n * m;
    for (; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t / m), j = (int)(t % m);
        out[t] = v[(long long)idx[i] * m + j];
    }
}

}  // namespace mgb200
