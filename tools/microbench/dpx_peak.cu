// Micro-benchmark: issue rates of the integer/DPX instructions the DP kernels are built from.
// Standalone (nvcc -> executable). Prints warp-lane ops per clock per SM for each instruction mix.
// Used to set the compute-roofline denominator for the DP kernels.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

constexpr int CH = 8;      // independent chains per thread
constexpr int ITERS = 4096;

template <int MODE>
__global__ void __launch_bounds__(1024, 1) kern(unsigned* out, long long* cyc, unsigned one, unsigned seed) {
    unsigned a[CH], b[CH], c[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) { a[k] = seed + threadIdx.x * 7 + k; b[k] = seed * 3 + k * 11 + threadIdx.x; c[k] = k + 1; }
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int k = 0; k < CH; ++k) {
            if (MODE == 0) {            // VIMNMX3.U16x2 only
                a[k] = __vimax3_u16x2(a[k], b[k], c[k]);
                b[k] = __vimax3_u16x2(b[k], c[k], a[k]);
            } else if (MODE == 1) {     // IADD3 only
                asm volatile("add.u32 %0, %0, %1;" : "+r"(a[k]) : "r"(b[k]));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(b[k]) : "r"(c[k]));
            } else if (MODE == 2) {     // IMAD (fma pipe) only
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[k]) : "r"(one), "r"(b[k]));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[k]) : "r"(one), "r"(c[k]));
            } else if (MODE == 3) {     // IMAD + VIMNMX3.U16x2 (the DP cell)
                unsigned t;
                asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(t) : "r"(b[k]), "r"(one), "r"(c[k]));
                a[k] = __vimax3_u16x2(a[k], b[k], t);
                b[k] = t ^ 0u;
            } else if (MODE == 4) {     // IADD + VIMNMX3.U16x2
                unsigned t = b[k] + c[k];
                a[k] = __vimax3_u16x2(a[k], b[k], t);
                b[k] = t;
            } else if (MODE == 5) {     // VIMNMX.U16x2 + VIADDMNMX.U16x2
                unsigned t = __vmaxu2(a[k], b[k]);
                a[k] = __viaddmax_u16x2(b[k], c[k], t);
                b[k] = t;
            } else if (MODE == 6) {     // VIADDMNMX.U16x2 only
                a[k] = __viaddmax_u16x2(a[k], c[k], b[k]);
                b[k] = __viaddmax_u16x2(b[k], c[k], a[k]);
            } else if (MODE == 7) {     // VIMNMX3 s32
                a[k] = __vimax3_s32(a[k], b[k], c[k]);
                b[k] = __vimax3_s32(b[k], c[k], a[k]);
            } else if (MODE == 8) {     // LOP3 only
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[k]) : "r"(b[k]), "r"(c[k]));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b[k]) : "r"(c[k]), "r"(a[k]));
            } else if (MODE == 9) {     // IMAD + IMAD + VIMNMX3 : 2 fma-pipe : 1 alu
                unsigned t, u;
                asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(t) : "r"(b[k]), "r"(one), "r"(c[k]));
                asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(u) : "r"(a[k]), "r"(one), "r"(c[k]));
                a[k] = __vimax3_u16x2(u, b[k], t);
                b[k] = t;
            }
        }
    }
    long long t1 = clock64();
    unsigned acc = 0;
#pragma unroll
    for (int k = 0; k < CH; ++k) acc ^= a[k] ^ b[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// shared-memory multicast: each lane reads one of NDIST distinct 16-byte chunks (LDS.128)
template <int NDIST, int STRIDE_WORDS>
__global__ void __launch_bounds__(1024, 1) lds_kern(unsigned* out, long long* cyc, unsigned seed) {
    __shared__ uint4 sm[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = make_uint4(i, i + 1, i + 2, i + 3);
    __syncthreads();
    unsigned sel = ((threadIdx.x * 2654435761u + seed) >> 7) % NDIST;
    const uint4* p = sm + sel * (STRIDE_WORDS / 4);
    unsigned acc = 0;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            uint4 v = p[k];
            acc += v.x ^ v.y ^ v.z ^ v.w;
        }
        p += (acc & 1) ? 0 : 0;
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int ops_per_chain_iter) {
    int nsm = 148;
    unsigned* out; long long* cyc;
    CK(cudaMalloc(&out, nsm * 1024 * 4)); CK(cudaMalloc(&cyc, nsm * 8));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<MODE><<<nsm, 1024>>>(out, cyc, 1u, 12345u);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    kern<MODE><<<nsm, 1024>>>(out, cyc, 1u, 12345u);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[148]; CK(cudaMemcpy(h, cyc, nsm * 8, cudaMemcpyDeviceToHost));
    double avg = 0; for (int i = 0; i < nsm; ++i) avg += h[i]; avg /= nsm;
    double lane_ops = 1024.0 * ITERS * CH * ops_per_chain_iter;
    printf("%-40s lane-ops/clk/SM = %7.2f  (cycles %.0f, %.3f ms, => %.0f MHz, %.2f T lane-ops/s chip)\n", name, lane_ops / avg, avg, ms,
           avg / (ms * 1e3), lane_ops * nsm / (ms * 1e-3) / 1e12);
    cudaFree(out); cudaFree(cyc);
}

template <int NDIST, int STRIDE>
void run_lds(const char* name) {
    int nsm = 148;
    unsigned* out; long long* cyc;
    CK(cudaMalloc(&out, nsm * 1024 * 4)); CK(cudaMalloc(&cyc, nsm * 8));
    lds_kern<NDIST, STRIDE><<<nsm, 1024>>>(out, cyc, 777u);
    CK(cudaDeviceSynchronize());
    lds_kern<NDIST, STRIDE><<<nsm, 1024>>>(out, cyc, 777u);
    CK(cudaDeviceSynchronize());
    long long h[148]; CK(cudaMemcpy(h, cyc, nsm * 8, cudaMemcpyDeviceToHost));
    double avg = 0; for (int i = 0; i < nsm; ++i) avg += h[i]; avg /= nsm;
    double warp_lds = 32.0 * ITERS * 8;
    printf("%-40s cycles per warp-LDS.128 per SM = %6.3f\n", name, avg / warp_lds);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("device %s, %d SMs, clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    run<0>("VIMNMX3.U16x2", 2);
    run<1>("IADD3", 2);
    run<2>("IMAD", 2);
    run<3>("IMAD + VIMNMX3.U16x2 (+LOP)", 2);
    run<4>("IADD + VIMNMX3.U16x2", 2);
    run<5>("VIMNMX.U16x2 + VIADDMNMX.U16x2", 2);
    run<6>("VIADDMNMX.U16x2", 2);
    run<7>("VIMNMX3 (s32)", 2);
    run<8>("LOP3", 2);
    run<9>("2 IMAD + VIMNMX3.U16x2", 3);
    run_lds<1, 64>("LDS.128 broadcast (1 addr)");
    run_lds<5, 100>("LDS.128 5 addrs, stride 100 words");
    run_lds<5, 36>("LDS.128 5 addrs, stride 36 words");
    run_lds<32, 4>("LDS.128 32 distinct consecutive");
    return 0;
}
