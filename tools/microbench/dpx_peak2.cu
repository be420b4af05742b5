// Second micro-benchmark: which pipe each DP instruction issues on (pairs of opcodes), and LDS.128 multicast cost.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
constexpr int CH = 8;
constexpr int ITERS = 2048;

#define OP_LOP3(d, x, y)   asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(d) : "r"(x), "r"(y))
#define OP_IMAD(d, x, y)   asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(d) : "r"(x), "r"(y))
#define OP_MAX3(d, x, y)   asm volatile("{.reg .b32 t; max.u16x2 t, %1, %2; max.u16x2 %0, %0, t;}" : "+r"(d) : "r"(x), "r"(y))
#define OP_MAX2(d, x)      asm volatile("max.u16x2 %0, %0, %1;" : "+r"(d) : "r"(x))
#define OP_PRMT(d, x, y)   asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(d) : "r"(x), "r"(y))
#define OP_SHF(d, x, y)    asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(d) : "r"(x), "r"(y))
#define OP_ADD(d, x)       asm volatile("add.u32 %0, %0, %1;" : "+r"(d) : "r"(x))
#define OP_ADDI(d)         asm volatile("add.u32 %0, %0, 0x00030003;" : "+r"(d))
#define OP_IMADI(d, x)     asm volatile("mad.lo.u32 %0, %0, 3, %1;" : "+r"(d) : "r"(x))

template <int MODE>
__global__ void __launch_bounds__(1024, 1) kern(unsigned* out, long long* cyc, unsigned one, unsigned seed) {
    unsigned a[CH], b[CH], c[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) { a[k] = seed + threadIdx.x * 7 + k; b[k] = seed * 3 + k * 11 + threadIdx.x; c[k] = one + k; }
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int k = 0; k < CH; ++k) {
            if (MODE == 0) { a[k] = __vimax3_u16x2(a[k], b[k], c[k]); OP_LOP3(b[k], c[k], one); }
            if (MODE == 1) { OP_IMAD(a[k], one, c[k]); OP_LOP3(b[k], c[k], one); }
            if (MODE == 2) { a[k] = __vimax3_u16x2(a[k], b[k], c[k]); OP_IMAD(b[k], one, c[k]); }
            if (MODE == 3) { OP_MAX2(a[k], b[k]); OP_MAX2(b[k], c[k]); }
            if (MODE == 4) { OP_MAX2(a[k], c[k]); OP_LOP3(b[k], c[k], one); }
            if (MODE == 5) { OP_MAX2(a[k], c[k]); OP_IMAD(b[k], one, c[k]); }
            if (MODE == 6) { a[k] = __vimax3_u16x2(a[k], b[k], b[k]); b[k] = __vimax3_u16x2(b[k], c[k], c[k]); }
            if (MODE == 7) { OP_PRMT(a[k], c[k], one); OP_IMAD(b[k], one, c[k]); }
            if (MODE == 8) { OP_ADD(a[k], c[k]); OP_ADD(b[k], c[k]); }
            if (MODE == 9) { OP_ADDI(a[k]); OP_ADDI(b[k]); }
            if (MODE == 10) { OP_IMADI(a[k], c[k]); OP_IMADI(b[k], c[k]); }
            if (MODE == 11) { a[k] = __viaddmax_u16x2(a[k], c[k], b[k]); OP_IMAD(b[k], one, c[k]); }
            if (MODE == 12) { a[k] = __vimax3_u16x2(a[k], b[k], c[k]); OP_ADD(b[k], c[k]); }
            if (MODE == 13) { a[k] = __vimax3_u16x2(a[k], b[k], c[k]); OP_ADDI(b[k]); }
            if (MODE == 14) { OP_MAX2(a[k], c[k]); OP_ADD(b[k], c[k]); }
            if (MODE == 15) { OP_MAX2(a[k], c[k]); OP_MAX2(a[k], b[k]); OP_ADD(b[k], c[k]); }
        }
    }
    long long t1 = clock64();
    unsigned acc = 0;
#pragma unroll
    for (int k = 0; k < CH; ++k) acc ^= a[k] ^ b[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int NDIST>
__global__ void __launch_bounds__(1024, 1) lds_kern(unsigned* out, long long* cyc, unsigned seed, int stride_bytes, int lane_mode) {
    __shared__ uint4 sm[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = make_uint4(i, i + 1, i + 2, i + 3);
    __syncthreads();
    unsigned sel = lane_mode == 0 ? (((threadIdx.x * 2654435761u + seed) >> 7) % NDIST) : (threadIdx.x % NDIST);
    unsigned addr = (unsigned)__cvta_generic_to_shared(sm) + sel * stride_bytes;
    unsigned acc = 0;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            unsigned x, y, z, w;
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(addr + k * 16));
            acc += x ^ y ^ z ^ w;
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int ops) {
    int nsm = 148; unsigned* out; long long* cyc;
    CK(cudaMalloc(&out, nsm * 1024 * 4)); CK(cudaMalloc(&cyc, nsm * 8));
    kern<MODE><<<nsm, 1024>>>(out, cyc, 1u, 12345u); CK(cudaDeviceSynchronize());
    kern<MODE><<<nsm, 1024>>>(out, cyc, 1u, 12345u); CK(cudaDeviceSynchronize());
    long long h[148]; CK(cudaMemcpy(h, cyc, nsm * 8, cudaMemcpyDeviceToHost));
    double avg = 0; for (int i = 0; i < nsm; ++i) avg += h[i]; avg /= nsm;
    printf("%-44s warp-inst/clk/SM = %5.2f  (lane-ops/clk/SM %6.1f)\n", name, 32.0 * ITERS * CH * ops / avg, 1024.0 * ITERS * CH * ops / avg);
    cudaFree(out); cudaFree(cyc);
}
template <int NDIST>
void run_lds(const char* name, int stride_bytes, int lane_mode) {
    int nsm = 148; unsigned* out; long long* cyc;
    CK(cudaMalloc(&out, nsm * 1024 * 4)); CK(cudaMalloc(&cyc, nsm * 8));
    lds_kern<NDIST><<<nsm, 1024>>>(out, cyc, 777u, stride_bytes, lane_mode); CK(cudaDeviceSynchronize());
    lds_kern<NDIST><<<nsm, 1024>>>(out, cyc, 777u, stride_bytes, lane_mode); CK(cudaDeviceSynchronize());
    long long h[148]; CK(cudaMemcpy(h, cyc, nsm * 8, cudaMemcpyDeviceToHost));
    double avg = 0; for (int i = 0; i < nsm; ++i) avg += h[i]; avg /= nsm;
    printf("%-44s cycles per warp-LDS.128 per SM = %6.3f\n", name, avg / (32.0 * ITERS * 8));
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<0>("VIMNMX3.U16x2 + LOP3", 2);
    run<1>("IMAD + LOP3", 2);
    run<2>("VIMNMX3.U16x2 + IMAD", 2);
    run<3>("VIMNMX.U16x2 (2-input) x2", 2);
    run<4>("VIMNMX.U16x2 + LOP3", 2);
    run<5>("VIMNMX.U16x2 + IMAD", 2);
    run<6>("VIMNMX3.U16x2 with 2 distinct regs", 2);
    run<7>("PRMT + IMAD", 2);
    run<8>("ADD reg x2", 2);
    run<9>("ADD imm x2", 2);
    run<10>("IMAD imm x2", 2);
    run<11>("VIADDMNMX.U16x2 + IMAD", 2);
    run<12>("VIMNMX3.U16x2 + ADD reg", 2);
    run<13>("VIMNMX3.U16x2 + ADD imm", 2);
    run<14>("VIMNMX.U16x2 + ADD reg", 2);
    run<15>("2x VIMNMX.U16x2 + ADD reg", 3);
    run_lds<1>("LDS.128 broadcast (1 addr)", 0, 0);
    run_lds<5>("LDS.128 5 addrs rand lanes, stride 400B", 400, 0);
    run_lds<5>("LDS.128 5 addrs rand lanes, stride 144B", 144, 0);
    run_lds<5>("LDS.128 5 addrs rand lanes, stride 128B", 128, 0);
    run_lds<32>("LDS.128 32 distinct consecutive", 16, 1);
    run_lds<32>("LDS.128 32 distinct stride 112B", 112, 1);
    run_lds<8>("LDS.128 8 addrs lane%8, stride 16B", 16, 1);
    return 0;
}
