// Micro-benchmark: per-SM throughput of the instructions the attention softmax can be built from.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu mufu.cu
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

constexpr int ITERS = 4096;
constexpr int UNROLL = 8;

template <int OP>
__global__ void __launch_bounds__(512) k(float* out, float seed) {
    float a[UNROLL];
    uint32_t u[UNROLL];
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) { a[i] = seed + threadIdx.x * 1e-3f + i; u[i] = __float_as_uint(a[i]); }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < UNROLL; ++i) {
            if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (OP == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u[i]));
            if (OP == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(u[i]));
            if (OP == 3) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %1;" : "=r"(u[i]) : "f"(__uint_as_float(u[i])));
            if (OP == 13) { asm volatile("cvt.rn.bf16x2.f32 %0, %1, %1;" : "=r"(u[i]) : "f"(__uint_as_float(u[i]))); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i])); }
            if (OP == 14) { asm volatile("prmt.b32 %0, %0, %0, 0x7632;" : "+r"(u[i])); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i])); }
            if (OP == 4) asm volatile("cvt.rn.f16x2.f32 %0, %1, %1;" : "=r"(u[i]) : "f"(__uint_as_float(u[i])));
            if (OP == 5) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a[i]));
            if (OP == 6) asm volatile("prmt.b32 %0, %0, %0, 0x7632;" : "+r"(u[i]));
            if (OP == 7) asm volatile("max.f32 %0, %0, %0, %0;" : "+f"(a[i]));
            if (OP == 8) asm volatile("fma.rn.f16x2 %0, %0, %0, %0;" : "+r"(u[i]));
            if (OP == 9) asm volatile("cvt.rz.bf16x2.f32 %0, %1, %1;" : "=r"(u[i]) : "f"(__uint_as_float(u[i])));
            if (OP == 10) asm volatile("ex2.approx.f16 %0, %0;" : "+h"(*(unsigned short*)&u[i]));
            if (OP == 11) asm volatile("add.f32 %0, %0, %0;" : "+f"(a[i]));
            if (OP == 12) asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(*(unsigned long long*)&a[i & ~1]));
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) s += a[i] + __uint_as_float(u[i]);
    if (s == 12345.678f) out[0] = s;
}

template <int OP>
void run(const char* name, int per_instr, int threads = 512) {
    float* d; cudaMalloc(&d, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int sms = 148;
    k<OP><<<sms, threads>>>(d, 1.0f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<OP><<<sms, threads>>>(d, 1.0f);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    double instr = (double)ITERS * UNROLL * threads;   // thread-instructions per SM
    double cyc = ms * 1e-3 * clk_khz * 1e3;
    printf("%-28s T=%3d %8.3f ms  %7.2f thread-instr/clk/SM (at %d MHz nominal)  -> %7.2f elems/clk/SM  err=%s\n", name, threads, ms, instr / cyc, clk_khz / 1000,
           per_instr * instr / cyc, cudaGetErrorString(cudaGetLastError()));
    cudaFree(d);
}

int main() {
    run<0>("ex2.approx.ftz.f32", 1);
    run<0>("ex2.approx.ftz.f32", 1, 128);
    run<0>("ex2.approx.ftz.f32", 1, 256);
    run<5>("fma.rn.f32", 1, 128);
    run<1>("ex2.approx.f16x2", 2);
    run<2>("ex2.approx.ftz.bf16x2", 2);
    run<10>("ex2.approx.f16", 1);
    run<3>("cvt.rn.bf16x2.f32", 2);
    run<9>("cvt.rz.bf16x2.f32", 2);
    run<4>("cvt.rn.f16x2.f32", 2);
    run<13>("cvt.rn.bf16x2 + ex2 (pairs)", 1);
    run<14>("prmt + ex2 (pairs)", 1);
    run<5>("fma.rn.f32", 1);
    run<12>("fma.rn.f32x2", 2);
    run<11>("add.f32", 1);
    run<6>("prmt.b32", 1);
    run<7>("max3.f32", 1);
    run<8>("fma.rn.f16x2", 2);
    return 0;
}
