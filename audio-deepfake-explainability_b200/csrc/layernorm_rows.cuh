// LayerNorm of RB rows held by one warp (one warp per row, lane l owns the float4 vectors l, l + 32, l + 64 of the row).
//
// ONE definition of the arithmetic for every kernel that normalises (layernorm_kernel in reduce_kernels.cu and the
// LayerNorm tail of the residual GEMM in gemm_tcgen05.cu), so that "LayerNorm as its own pass" and "LayerNorm fused behind
// the GEMM" give the same bits by construction.  Replaces nn.LayerNorm of the third-party encoder block reached through
// LocalSonnics.predict (src/sonics_api.py:259-271).
//
// Packed fp32x2 math (FADD2 / FMUL2 / FFMA2: two IEEE round-to-nearest results per issue slot) and a multiplication by 1/D
// instead of two IEEE divisions per row: ~75 instructions per row instead of ~160, which is what lets eight warps keep up
// with a residual GEMM.  Every step runs over all RB rows before the next step starts, so the shuffle and rsqrt latencies of
// the rows overlap (written row after row, ptxas serialises the dependent chains).
//   s    = sum_i [(x_i + z_i) + (y_i + w_i)] accumulated pairwise, then butterfly over the lanes (xor 16, 8, 4, 2, 1)
//   mean = s * (1/D);  d = v - mean;  q = sum d^2 (same order);  rstd = rsqrt(q * (1/D) + eps)
//   out  = (d * rstd) * gamma + beta, rounded to bf16 (round to nearest even)
#pragma once
#include "ptx.cuh"

namespace b200x {

struct LnPair2 { uint64_t xy, zw; };   // a float4 as two packed fp32x2 values

template <int RB, typename Store>
__device__ __forceinline__ void layernorm_rows(const float4 (&v)[RB][3], int vpl, float inv_d, float eps, const float* __restrict__ gamma,
                                               const float* __restrict__ beta, int lane, Store&& store) {
    LnPair2 d[RB][3];
    float s[RB];
#pragma unroll
    for (int r = 0; r < RB; ++r) {
        uint64_t acc = 0ull;                                         // (+0, +0)
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            d[r][i].xy = pack_f32x2(v[r][i].x, v[r][i].y);
            d[r][i].zw = pack_f32x2(v[r][i].z, v[r][i].w);
            if (i < vpl) acc = fadd2(acc, fadd2(d[r][i].xy, d[r][i].zw));
        }
        float lo, hi;
        unpack_f32x2(acc, lo, hi);
        s[r] = lo + hi;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int r = 0; r < RB; ++r) s[r] += __shfl_xor_sync(0xffffffffu, s[r], o);
    }
#pragma unroll
    for (int r = 0; r < RB; ++r) {
        const float nm = -(s[r] * inv_d);
        const uint64_t nm2 = pack_f32x2(nm, nm);
        uint64_t acc = 0ull;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            if (i < vpl) {
                d[r][i].xy = fadd2(d[r][i].xy, nm2);
                d[r][i].zw = fadd2(d[r][i].zw, nm2);
                acc = ffma2(d[r][i].xy, d[r][i].xy, acc);
                acc = ffma2(d[r][i].zw, d[r][i].zw, acc);
            }
        }
        float lo, hi;
        unpack_f32x2(acc, lo, hi);
        s[r] = lo + hi;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int r = 0; r < RB; ++r) s[r] += __shfl_xor_sync(0xffffffffu, s[r], o);
    }
#pragma unroll
    for (int r = 0; r < RB; ++r) s[r] = rsqrtf(fmaf(s[r], inv_d, eps));
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        if (i < vpl) {
            const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * i);
            const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + lane + 32 * i);
            const uint64_t gxy = pack_f32x2(g.x, g.y), gzw = pack_f32x2(g.z, g.w);
            const uint64_t bxy = pack_f32x2(b.x, b.y), bzw = pack_f32x2(b.z, b.w);
#pragma unroll
            for (int r = 0; r < RB; ++r) {
                const uint64_t rs2 = pack_f32x2(s[r], s[r]);
                float ox, oy, oz, ow;
                unpack_f32x2(ffma2(fmul2(d[r][i].xy, rs2), gxy, bxy), ox, oy);
                unpack_f32x2(ffma2(fmul2(d[r][i].zw, rs2), gzw, bzw), oz, ow);
                store(r, i, make_float4(ox, oy, oz, ow));
            }
        }
    }
}

}  // namespace b200x
