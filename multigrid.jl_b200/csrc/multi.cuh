// Single-process multi-GPU entry (SURVEY.md section 8(b) "Threading"): ONE handle drives G devices, so a single
// `ccall` from one Julia task is enough - no mpirun / torchrun / Distributed workers on the product path.
//
// A MultiHandle owns one ordinary hierarchy handle per device, each bound to rank g of an NCCL communicator created
// inside the process.  Every entry point takes GLOBAL arrays exactly as the reference holds them (the CSC arrays of
// As[l], Ps[l], Rs[l], the full b and x), slices the contiguous row ranges of the z-slab partition
// (getOriginalBoundingBoxCells with NumCells = [1,1,G], DDIndices.jl:41-47: the caller passes the row offsets) and
// runs the per-rank call of every device on its own host thread; the collective steps inside those calls (ghost
// planning, the NVLink peer-memory halo exchange of p2p.cuh, Krylov all-reduces) meet as they do between processes.
// Peers of the same process map each other's receive buffers by plain peer access instead of CUDA IPC.
#pragma once
#include <exception>
#include <memory>
#include <string>
#include <functional>
#include <thread>
#include <vector>

namespace mgb200 {

struct MultiHandle {
    int G = 0;
    int val_type = 0, levels = 0, nrhs = 1;
    std::vector<int> devices;
    std::vector<mgb200_handle> h;
    std::vector<int64_t> fine_offsets;          // G + 1 global row offsets of level 1
    std::string error;
};

// run f(g) for every rank on its own thread; the first failure (status != 0) is reported
static int multi_run(MultiHandle* M, const std::function<int(int)>& f) {
    std::vector<int> st(M->G, 0);
    std::vector<std::string> msg(M->G);
    std::vector<std::thread> th;
    for (int g = 0; g < M->G; ++g)
        th.emplace_back([&, g] {
            st[g] = f(g);
            if (st[g] != 0) msg[g] = mgb200_last_error();
        });
    for (auto& t : th) t.join();
    for (int g = 0; g < M->G; ++g)
        if (st[g] != 0) {
            M->error = "rank " + std::to_string(g) + ": " + msg[g];
            return st[g];
        }
    return 0;
}

// columns [c0, c1) of a CSC matrix as a CSC block with its own column pointers (same index base)
struct CscSlice {
    std::vector<int64_t> colptr;
    const int64_t* rowval;
    const unsigned char* nzval;
};
static CscSlice csc_columns(const int64_t* colptr, const int64_t* rowval, const void* nzval, size_t val_bytes, int64_t c0,
                            int64_t c1, int base) {
    CscSlice S;
    const int64_t k0 = colptr[c0] - base;
    S.colptr.resize(c1 - c0 + 1);
    for (int64_t c = c0; c <= c1; ++c) S.colptr[c - c0] = colptr[c] - k0;
    S.rowval = rowval + k0;
    S.nzval = static_cast<const unsigned char*>(nzval) + (size_t)k0 * val_bytes;
    return S;
}

}  // namespace mgb200
