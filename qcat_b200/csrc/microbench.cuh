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

// MODE 0: the add as a real multiply-add with a run-time multiplier of 1 (always an IMAD, three register operands);
// MODE 1: the add written with a literal 1 -- ptxas then emits its own mix of IMAD.IADD and IADD3, as in the kernels;
// MODE 2: MODE 0 with the addend equal to the multiplier register in one chain (tools/microbench/dpx_peak2.cu's form).
// Register assignment moves the measured rate by a few percent (operand bank conflicts), so the reported peak is the
// fastest of the variants (MODE 3 below included): the most demanding denominator.
template <int MODE>
__global__ void __launch_bounds__(1024, 1) k_microbench_cell(unsigned *out, long long *cycles, unsigned seed, unsigned one)
{
    unsigned a[kMbChains], b[kMbChains], c[kMbChains];
#pragma unroll
    for (int k = 0; k < kMbChains; ++k) {
        a[k] = seed + threadIdx.x * 7 + k; b[k] = seed * 3 + k * 11 + threadIdx.x;
        c[k] = MODE == 2 ? one + k : one * 0x00010001u + k;      // run-time values: register operands, like the kernels'
    }
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < kMbIters; ++it) {
#pragma unroll
        for (int k = 0; k < kMbChains; ++k) {
            a[k] = __vimax3_u16x2(a[k], b[k], c[k]);                                             // max(left, up, diag term)
            if (MODE == 1) asm volatile("mad.lo.u32 %0, %0, 1, %1;" : "+r"(b[k]) : "r"(c[k]));    // diag + substitution score
            else asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[k]) : "r"(one), "r"(c[k]));
        }
    }
    long long t1 = clock64();
    unsigned acc = 0;
#pragma unroll
    for (int k = 0; k < kMbChains; ++k) acc ^= a[k] ^ b[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// MODE 3: the kernel of tools/microbench/dpx_peak2.cu (mode "VIMNMX3.U16x2 + IMAD") as it stands there -- same
// signature, same statements -- so that ptxas assigns the same registers as in the stand-alone tool
// (profiles/dpx_peak2_r01.txt: 2.63 warp-instructions per clock per SM).
__global__ void __launch_bounds__(1024, 1) k_microbench_cell_ref(unsigned *out, long long *cyc, unsigned one, unsigned seed)
{
    unsigned a[kMbChains], b[kMbChains], c[kMbChains];
#pragma unroll
    for (int k = 0; k < kMbChains; ++k) { a[k] = seed + threadIdx.x * 7 + k; b[k] = seed * 3 + k * 11 + threadIdx.x; c[k] = one + k; }
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < kMbIters; ++it) {
#pragma unroll
        for (int k = 0; k < kMbChains; ++k) {
            a[k] = __vimax3_u16x2(a[k], b[k], c[k]);
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[k]) : "r"(one), "r"(c[k]));
        }
    }
    long long t1 = clock64();
    unsigned acc = 0;
#pragma unroll
    for (int k = 0; k < kMbChains; ++k) acc ^= a[k] ^ b[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
inline int microbench_one(int nsm, unsigned *out, long long *cyc, double *cells_per_clk, double *mhz, std::string &err)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best_cpc = 0.0, best_mhz = 0.0;
    std::vector<long long> h(nsm);
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        if (MODE == 3) k_microbench_cell_ref<<<nsm, 1024>>>(out, cyc, 1u, 12345u);
        else k_microbench_cell<MODE == 3 ? 0 : MODE><<<nsm, 1024>>>(out, cyc, 12345u + rep, 1u);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { err = "microbench kernel failed"; cudaEventDestroy(e0); cudaEventDestroy(e1); return 1; }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        cudaMemcpy(h.data(), cyc, (size_t)nsm * 8, cudaMemcpyDeviceToHost);
        double avg = 0; for (long long v : h) avg += (double)v; avg /= nsm;
        const double cpc = 2.0 * 1024.0 * kMbIters * kMbChains / avg;      // packed 16-bit cells per clock per SM
        if (rep > 0 && cpc > best_cpc) { best_cpc = cpc; best_mhz = avg / (ms * 1e3); }
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *cells_per_clk = best_cpc; *mhz = best_mhz;
    return 0;
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
    double cpc[4] = {0, 0, 0, 0}, mhz[4] = {0, 0, 0, 0};
    int rc = microbench_one<0>(nsm, out, cyc, &cpc[0], &mhz[0], err);
    if (!rc) rc = microbench_one<1>(nsm, out, cyc, &cpc[1], &mhz[1], err);
    if (!rc) rc = microbench_one<2>(nsm, out, cyc, &cpc[2], &mhz[2], err);
    if (!rc) rc = microbench_one<3>(nsm, out, cyc, &cpc[3], &mhz[3], err);
    cudaFree(out); cudaFree(cyc);
    if (rc) return rc;
    int best = 0;
    for (int m = 1; m < 4; ++m) if (cpc[m] > cpc[best]) best = m;
    *sm_mhz = mhz[best];
    *cells_per_second = cpc[best] * nsm * mhz[best] * 1e6;     // per-clock rate of the fastest variant at the clock it ran at
    return 0;
}

}  // namespace qcb
