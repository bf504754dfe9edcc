// Micro-benchmark: tcgen05.mma issue-to-retire throughput per SM for the shapes / operand sources the attention and GEMM
// kernels use.  build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../audio-deepfake-explainability_b200/csrc -o umma umma.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace b200x;

struct Cfg { int n; int ts; int b_mn; int alt_d; int iters; };

__global__ void __launch_bounds__(128, 1) k(Cfg c, long long* cycles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    fence_proxy_async();
    if (warp == 0) tmem_alloc<512>(&slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = slot;
    if (warp == 1 && elect_one()) {
        const uint32_t idesc = make_idesc_bf16(128, c.n, c.b_mn != 0);
        const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 65536);
        long long t0 = clock64();
        for (int it = 0; it < c.iters; ++it) {
            // one "tile": 8 K-steps over 128-wide K (4 per 64-wide swizzle row for K-major; 8 for the MN-major V layout)
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
                const uint32_t d = tb + ((c.alt_d && (it & 1)) ? 256 : 0);
                uint64_t bd;
                if (c.b_mn) bd = make_smem_desc_sw128(b_addr + ks * 2048, 16384, 1024);
                else bd = make_smem_desc_sw128(b_addr + (ks & 3) * 32 + (ks >> 2) * 32768, 16, 1024);
                if (c.ts) umma_ts(d, tb + 384 + ks * 8, bd, idesc, (ks != 0) ? 1u : 0u);
                else umma_ss(d, make_smem_desc_sw128(a_addr + (ks & 3) * 32 + (ks >> 2) * 16384, 16, 1024), bd, idesc, (ks != 0) ? 1u : 0u);
            }
        }
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        long long t1 = clock64();
        cycles[blockIdx.x] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tb);
}

void run(const char* name, Cfg c, int ctas = 148) {
    long long* d; cudaMalloc(&d, 148 * 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    k<<<ctas, 128, 200 * 1024>>>(c, d);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<<<ctas, 128, 200 * 1024>>>(c, d);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[148]; cudaMemcpy(h, d, ctas * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < ctas; ++i) avg += h[i]; avg /= ctas;
    const double per_mma = avg / (c.iters * 8.0);
    const double flops = 2.0 * 128 * c.n * 16 * 8.0 * c.iters * ctas;
    printf("%-44s ctas=%3d  %7.1f clk/MMA (ideal %5.1f)  %8.1f TFLOP/s (event)  err=%s\n", name, ctas, per_mma, 128.0 * c.n / 256.0,
           flops / (ms * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
    cudaFree(d);
}

int main() {
    const int it = 2000;
    run("SS N=128 K-major B (S = Q K^T)", {128, 0, 0, 0, it});
    run("SS N=128 K-major B, alternating D", {128, 0, 0, 1, it});
    run("SS N=256 K-major B", {256, 0, 0, 0, it});
    run("SS N=64  K-major B", {64, 0, 0, 0, it});
    run("SS N=64  MN-major B", {64, 0, 1, 0, it});
    run("TS N=64  MN-major B (O += P V)", {64, 1, 1, 0, it});
    run("TS N=64  MN-major B, alternating D", {64, 1, 1, 1, it});
    run("TS N=64  K-major B", {64, 1, 0, 0, it});
    run("TS N=128 K-major B", {128, 1, 0, 0, it});
    run("TS N=128 MN-major B", {128, 1, 1, 0, it});
    run("TS N=256 K-major B", {256, 1, 0, 0, it});
    run("SS N=128 K-major B, 1 CTA", {128, 0, 0, 0, it}, 1);
    run("TS N=64  MN-major B, 1 CTA", {64, 1, 1, 0, it}, 1);
    return 0;
}
