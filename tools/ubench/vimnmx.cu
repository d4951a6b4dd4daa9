// Throughput of the integer min/max flavours on sm_100a (per SM, per clock).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int OP>
__global__ void k(uint32_t* out, int iters, uint32_t seed) {
    uint32_t a[8];
    for (int i = 0; i < 8; ++i) a[i] = seed * (threadIdx.x + 1) + i * 0x10001u;
    uint32_t b = seed ^ 0x1234567u, c = seed + 0x7654321u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) asm volatile("max.u16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
            if (OP == 1) { asm volatile("{.reg .b32 t; max.u16x2 t, %0, %1; max.u16x2 %0, t, %2;}" : "+r"(a[i]) : "r"(b), "r"(c)); }
            if (OP == 2) asm volatile("max.s32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
            if (OP == 3) { asm volatile("{.reg .b32 t; max.s32 t, %0, %1; max.s32 %0, t, %2;}" : "+r"(a[i]) : "r"(b), "r"(c)); }
            if (OP == 4) asm volatile("max.s16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
            if (OP == 5) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
            if (OP == 6) { asm volatile("{.reg .b32 t; max.u32 t, %0, %1; max.u32 %0, t, %2;}" : "+r"(a[i]) : "r"(b), "r"(c)); }
        }
        b += 3; c ^= b;
    }
    uint32_t s = 0;
    for (int i = 0; i < 8; ++i) s ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
void run(const char* name) {
    uint32_t* d; cudaMalloc(&d, 148 * 1024 * 4);
    int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<148, 1024>>>(d, 100, 1);
    cudaEventRecord(e0);
    k<OP><<<148, 1024>>>(d, iters, 1);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double insts = 8.0 * iters * 1024;           // thread-instructions per SM
    double cyc = ms * 1e-3 * clk * 1e3;
    printf("%-22s %8.3f ms  %6.1f lane-ops/clk/SM (at %d MHz nominal)\n", name, ms, insts / cyc, clk / 1000);
    cudaFree(d);
}

int main() {
    run<0>("max.u16x2 (2-in)");
    run<1>("max.u16x2 x2 (3-in)");
    run<4>("max.s16x2 (2-in)");
    run<2>("max.s32 (2-in)");
    run<3>("max.s32 x2 (3-in)");
    run<6>("max.u32 x2 (3-in)");
    run<5>("add.u32");
    return 0;
}
