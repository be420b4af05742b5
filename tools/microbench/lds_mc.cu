// Micro-benchmark: cost of shared-memory loads whose lanes hit only a few distinct addresses (multicast),
// for 32/64/128-bit widths -- decides how the substitution profile is laid out for the packed DP kernel.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
constexpr int ITERS = 2048;

template <int WIDTH>
__global__ void __launch_bounds__(1024, 1) lds_kern(unsigned* out, long long* cyc, int ndist, int stride_bytes, int lane_mode, unsigned zero_mask) {
    __shared__ uint4 sm[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = make_uint4(i, i + 1, i + 2, i + 3);
    __syncthreads();
    unsigned lane = threadIdx.x & 31;
    unsigned sel = lane_mode == 0 ? ((lane * 2654435761u) >> 9) % ndist : (lane_mode == 1 ? lane % ndist : lane / (32 / ndist));
    unsigned addr = (unsigned)__cvta_generic_to_shared(sm) + sel * stride_bytes;
    unsigned acc = 0;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        unsigned a2 = addr + (acc & zero_mask);
        for (int k = 0; k < 8; ++k) {
            unsigned x = 0, y = 0, z = 0, w = 0;
            if (WIDTH == 16) asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(a2 + k * 16) : "memory");
            if (WIDTH == 8) asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(x), "=r"(y) : "r"(a2 + k * 8) : "memory");
            if (WIDTH == 4) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(x) : "r"(a2 + k * 4) : "memory");
            acc += x ^ y ^ z ^ w;
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int WIDTH>
void run(const char* name, int ndist, int stride_bytes, int lane_mode) {
    int nsm = 148; unsigned* out; long long* cyc;
    CK(cudaMalloc(&out, nsm * 1024 * 4)); CK(cudaMalloc(&cyc, nsm * 8));
    lds_kern<WIDTH><<<nsm, 1024>>>(out, cyc, ndist, stride_bytes, lane_mode, 0u); CK(cudaDeviceSynchronize());
    lds_kern<WIDTH><<<nsm, 1024>>>(out, cyc, ndist, stride_bytes, lane_mode, 0u); CK(cudaDeviceSynchronize());
    long long h[148]; CK(cudaMemcpy(h, cyc, nsm * 8, cudaMemcpyDeviceToHost));
    double avg = 0; for (int i = 0; i < nsm; ++i) avg += h[i]; avg /= nsm;
    printf("LDS.%-3d %-52s cycles per warp-load per SM = %6.3f\n", WIDTH * 8, name, avg / (32.0 * ITERS * 8));
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<4>("1 addr (broadcast)", 1, 0, 1);
    run<4>("32 distinct consecutive words", 32, 4, 1);
    run<4>("5 addrs, stride 112 B, random lanes", 5, 112, 0);
    run<4>("5 addrs, stride 128 B (same bank), random lanes", 5, 128, 0);
    run<8>("1 addr (broadcast)", 1, 0, 1);
    run<8>("32 distinct consecutive", 32, 8, 1);
    run<8>("5 addrs, stride 112 B, random lanes", 5, 112, 0);
    run<16>("1 addr (broadcast)", 1, 0, 1);
    run<16>("32 distinct consecutive", 32, 16, 1);
    run<16>("4 addrs, stride 112 B, random lanes", 4, 112, 0);
    run<16>("5 addrs, stride 112 B, random lanes", 5, 112, 0);
    run<16>("6 addrs, stride 112 B, random lanes", 6, 112, 0);
    run<16>("5 addrs, stride 112 B, lane%5", 5, 112, 1);
    run<16>("4 addrs, stride 112 B, lane/8 (one addr per quarter warp)", 4, 112, 2);
    run<16>("5 addrs, stride 128 B (same banks), random lanes", 5, 128, 0);
    run<16>("2 addrs, stride 112 B, random lanes", 2, 112, 0);
    run<16>("8 addrs, stride 16 B, lane%8", 8, 16, 1);
    return 0;
}
