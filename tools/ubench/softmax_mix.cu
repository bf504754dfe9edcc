// Micro-benchmark: the attention softmax instruction mix on registers only (no TMEM, no barriers):
// per 128 scores: 64 FMNMX3 + 64 FFMA2 + 128 MUFU.EX2 + 64 FADD2 + 64 F2FP.  Reports cycles per 128-score row pass.
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace b200x;

template <int MODE>
__device__ __forceinline__ void chunk(const uint32_t* r, uint32_t* pk, uint64_t c2, uint64_t nmc2, uint64_t zero2, uint64_t& acc_a, uint64_t& acc_b) {
    const uint64_t link_a = (MODE & 1) ? ffma2(acc_a, zero2, nmc2) : nmc2;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float x0, x1;
        unpack_f32x2(ffma2(pack_f32x2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), c2, link_a), x0, x1);
        float p0, p1;
        if (((MODE & 8) && (i & 3) == 3) || ((MODE & 16) && (i & 1))) exp2_poly2(x0, x1, p0, p1);
        else { p0 = (MODE & 2) ? x0 : ex2_approx(x0); p1 = (MODE & 2) ? x1 : ex2_approx(x1); }
        acc_a = fadd2(acc_a, pack_f32x2(p0, p1));
        pk[i] = pack_bf16(p0, p1);
    }
    const uint64_t link_b = (MODE & 1) ? ffma2(acc_b, zero2, nmc2) : nmc2;
#pragma unroll
    for (int i = 8; i < 16; ++i) {
        float x0, x1;
        unpack_f32x2(ffma2(pack_f32x2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), c2, link_b), x0, x1);
        float p0, p1;
        if (((MODE & 8) && (i & 3) == 3) || ((MODE & 16) && (i & 1))) exp2_poly2(x0, x1, p0, p1);
        else { p0 = (MODE & 2) ? x0 : ex2_approx(x0); p1 = (MODE & 2) ? x1 : ex2_approx(x1); }
        acc_b = fadd2(acc_b, pack_f32x2(p0, p1));
        pk[i] = pack_bf16(p0, p1);
    }
}

template <int MODE>
__global__ void __launch_bounds__(256, 1) k(int iters, float zero, float cscale, long long* cycles, uint32_t* sink) {
    uint32_t r[128];
#pragma unroll
    for (int i = 0; i < 128; ++i) r[i] = __float_as_uint(-0.01f * (i + threadIdx.x % 7));
    const uint64_t zero2 = pack_f32x2(zero, zero), c2 = pack_f32x2(cscale, cscale);
    uint64_t la = 0, lb = 0;
    uint32_t keep = 0;
    float m_ref = 0.f;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
        if (!(MODE & 4)) {
#pragma unroll
            for (int i = 0; i < 128; i += 8) {
                m0 = fmax3(m0, __uint_as_float(r[i]), __uint_as_float(r[i + 1]));
                m1 = fmax3(m1, __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
                m2 = fmax3(m2, __uint_as_float(r[i + 4]), __uint_as_float(r[i + 5]));
                m3 = fmax3(m3, __uint_as_float(r[i + 6]), __uint_as_float(r[i + 7]));
            }
            m_ref = fmaxf(m_ref, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));
        }
        const float mc = m_ref * cscale;
        const uint64_t nmc2 = pack_f32x2(-mc, -mc);
        uint32_t pk[64];
        chunk<MODE>(r, pk, c2, nmc2, zero2, la, lb);
        chunk<MODE>(r + 32, pk + 16, c2, nmc2, zero2, la, lb);
        chunk<MODE>(r + 64, pk + 32, c2, nmc2, zero2, la, lb);
        chunk<MODE>(r + 96, pk + 48, c2, nmc2, zero2, la, lb);
        // fold the outputs back into the inputs so that nothing is dead and every iteration depends on the previous one
#pragma unroll
        for (int i = 0; i < 64; ++i) { keep ^= pk[i]; }
        r[it & 127 ? 5 : 6] ^= (keep & 1u);
    }
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) cycles[blockIdx.x * 8 + (threadIdx.x >> 5)] = t1 - t0;
    float a, b; unpack_f32x2(fadd2(la, lb), a, b);
    if (a + b == 1234.5f) sink[0] = keep;
}

template <int MODE>
void run(const char* name, int threads) {
    long long* d; uint32_t* s; cudaMalloc(&d, 148 * 8 * 8); cudaMalloc(&s, 4);
    const int iters = 2000;
    k<MODE><<<148, threads>>>(iters, 0.f, 0.18f, d, s);
    cudaDeviceSynchronize();
    long long h[148 * 8]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; int n = 0;
    for (int b = 0; b < 148; ++b) for (int w = 0; w < threads / 32; ++w) { avg += h[b * 8 + w]; ++n; }
    printf("%-52s warps/SMSP=%d  %8.1f clk per 128-score row pass  err=%s\n", name, threads / 128, avg / n / iters, cudaGetErrorString(cudaGetLastError()));
    cudaFree(d); cudaFree(s);
}

int main() {
    run<0>("max + exp (MUFU), no links", 128);
    run<0>("max + exp (MUFU), no links", 256);
    run<1>("max + exp (MUFU), linked halves", 128);
    run<1>("max + exp (MUFU), linked halves", 256);
    run<8>("max + exp, 25% polynomial", 128);
    run<8>("max + exp, 25% polynomial", 256);
    run<9>("max + exp, 25% polynomial, linked", 256);
    run<16>("max + exp, 50% polynomial", 128);
    run<16>("max + exp, 50% polynomial", 256);
    run<2>("max + no MUFU", 128);
    run<2>("max + no MUFU", 256);
    run<4>("exp only (no max pass)", 128);
    run<4>("exp only (no max pass)", 256);
    return 0;
}
