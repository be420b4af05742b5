// Issue-rate micro-benchmark of the packed DP cell: the instruction pair the packed kernels execute per two cells --
// VIMNMX3.U16x2 and the diagonal add, which ptxas emits as IMAD.IADD (an IMAD with multiplier 1, so that it issues on
// the FMA pipe next to the ALU pipe's VIMNMX3; profiles/sass_k_barcode_fast_r02.txt).  The add is written here as a real
// multiply-add with a run-time multiplier of 1 so that it stays an IMAD.  This is the compute-roofline denominator
// reported next to the HBM one; tools/microbench/dpx_peak2.cu measures the other pairs.
#pragma once

#include <string>
#include <vector>
#include <cuda_runtime.h>

namespace qcb {

constexpr int kMbChains = 8;
constexpr int kMbIters = 2048;

__global__ void __launch_bounds__(1024, 1) k_microbench_cell(unsigned *out, long long *cycles, unsigned seed, unsigned one)
{
    unsigned a[kMbChains], b[kMbChains], c[kMbChains];
#pragma unroll
    for (int k = 0; k < kMbChains; ++k) { a[k] = seed + threadIdx.x * 7 + k; b[k] = seed * 3 + k * 11 + threadIdx.x; c[k] = one * 0x00010001u + k; }   // run-time values: register operands, like the kernels'
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < kMbIters; ++it) {
#pragma unroll
        for (int k = 0; k < kMbChains; ++k) {
            a[k] = __vimax3_u16x2(a[k], b[k], c[k]);                                         // max(left, up, diag term)
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[k]) : "r"(one), "r"(c[k]));    // diag + substitution score
        }
    }
    long long t1 = clock64();
    unsigned acc = 0;
#pragma unroll
    for (int k = 0; k < kMbChains; ++k) acc ^= a[k] ^ b[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

inline int microbench_cell_rate(double *cells_per_second, double *sm_mhz, std::string &err)
{
    int dev = 0, nsm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    unsigned *out = nullptr; long long *cyc = nullptr;
    if (cudaMalloc(&out, (size_t)nsm * 1024 * 4) != cudaSuccess || cudaMalloc(&cyc, (size_t)nsm * 8) != cudaSuccess) {
        err = "microbench allocation failed"; cudaFree(out); cudaFree(cyc); return 1;
    }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best_ms = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k_microbench_cell<<<nsm, 1024>>>(out, cyc, 12345u + rep, 1u);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { err = "microbench kernel failed"; cudaFree(out); cudaFree(cyc); return 1; }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best_ms) best_ms = ms;
    }
    std::vector<long long> h(nsm);
    cudaMemcpy(h.data(), cyc, (size_t)nsm * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (long long v : h) avg += (double)v; avg /= nsm;
    double cells = 2.0 * 1024.0 * kMbIters * kMbChains * nsm;      // two 16-bit cells per packed op pair
    *cells_per_second = cells / (best_ms * 1e-3);
    *sm_mhz = avg / (best_ms * 1e3);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(out); cudaFree(cyc);
    return 0;
}

}  // namespace qcb
