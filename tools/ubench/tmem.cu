// Micro-benchmark: tcgen05.ld / tcgen05.st throughput per SM.  build like umma.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace b200x;

template <int MODE>
__global__ void __launch_bounds__(256, 1) k(int iters, long long* cycles, float* sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc<512>(&slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16) + (warp >> 2) * 256;
    uint32_t r[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) r[i] = i;
    tmem_st16(tb, r); tmem_st16(tb + 16, r + 16);
    tmem_wait_st();
    __syncthreads();
    float acc = 0.f;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {            // 4 x ld32 then one wait (128 columns)
            uint32_t a[32], b[32], c[32], d[32];
            tmem_ld32(tb, a); tmem_ld32(tb + 32, b); tmem_ld32(tb + 64, c); tmem_ld32(tb + 96, d);
            tmem_wait_ld();
            acc += __uint_as_float(a[0] ^ b[7] ^ c[15] ^ d[31] ^ a[31] ^ b[0] ^ c[0] ^ d[0]);
        } else if (MODE == 1) {     // ld32 + wait each
            uint32_t a[32];
            tmem_ld32(tb + (it & 3) * 32, a);
            tmem_wait_ld();
            acc += __uint_as_float(a[0] ^ a[31] ^ a[13]);
        } else if (MODE == 2) {     // st16 x2 (32 columns) + wait
            tmem_st16(tb + (it & 3) * 32, r); tmem_st16(tb + (it & 3) * 32 + 16, r + 16);
            tmem_wait_st();
        }
    }
    long long t1 = clock64();
    if (threadIdx.x % 32 == 0) cycles[blockIdx.x * 8 + warp] = t1 - t0;
    if (acc == 1234.5f) sink[0] = acc;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(slot);
}

template <int MODE>
void run(const char* name, int threads, int cols_per_iter) {
    long long* d; float* s; cudaMalloc(&d, 148 * 8 * 8); cudaMalloc(&s, 4);
    const int iters = 4000;
    k<MODE><<<148, threads>>>(iters, d, s);
    cudaDeviceSynchronize();
    long long h[148 * 8]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; int n = 0;
    for (int b = 0; b < 148; ++b) for (int w = 0; w < threads / 32; ++w) { avg += h[b * 8 + w]; ++n; }
    avg /= n;
    const double bytes = (double)cols_per_iter * 4 * 32 * (threads / 32) * iters;     // per SM
    printf("%-40s warps=%d  %8.1f clk/iter  %7.1f B/clk/SM  err=%s\n", name, threads / 32, avg / iters, bytes / avg, cudaGetErrorString(cudaGetLastError()));
    cudaFree(d); cudaFree(s);
}

int main() {
    run<0>("4 x ld.32x32b.x32 + wait (128 cols)", 128, 128);
    run<0>("4 x ld.32x32b.x32 + wait (128 cols)", 256, 128);
    run<1>("ld.32x32b.x32 + wait (32 cols)", 128, 32);
    run<1>("ld.32x32b.x32 + wait (32 cols)", 256, 32);
    run<2>("2 x st.32x32b.x16 + wait (32 cols)", 128, 32);
    run<2>("2 x st.32x32b.x16 + wait (32 cols)", 256, 32);
    return 0;
}
