// FP64 issue-rate probe: how many DFMA / DMUL+DADD per clock per SM does this GPU sustain?
#include <cuda_runtime.h>
#include <cstdio>
template <int MODE>
__global__ void k(double* out, int iters, double a, double b) {
    double acc[8];
    for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) acc[i] = fma(acc[i], a, b);
            else if (MODE == 1) acc[i] = __dadd_rn(__dmul_rn(acc[i], a), b);
            else acc[i] = __dadd_rn(acc[i], b);
        }
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    int nsm; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double* out; cudaMalloc(&out, (size_t)nsm * 8 * 256 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int mode = 0; mode < 3; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<nsm * 8, 256>>>(out, iters, 1.0000001, 1e-9);
            else if (mode == 1) k<1><<<nsm * 8, 256>>>(out, iters, 1.0000001, 1e-9);
            else k<2><<<nsm * 8, 256>>>(out, iters, 1.0000001, 1e-9);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double instr = (double)nsm * 8 * 256 * iters * 8 * (mode == 1 ? 2 : 1);
        printf("mode %d (%s): %.3f ms, %.2f G thread-instr/s, %.2f lanes/clk/SM (clock %d kHz)\n", mode,
               mode == 0 ? "DFMA" : (mode == 1 ? "DMUL+DADD" : "DADD"), ms, instr / ms / 1e6,
               instr / (ms * 1e-3) / nsm / (clk * 1e3), clk);
    }
    return 0;
}
