// Warp-level 1024-point complex FFT (the engine's 2048-point real STFT / iSTFT frames are packed into it).
//
// N = 32 x 32 Cooley-Tukey: every lane runs a 32-point FFT entirely in registers, the warp transposes once
// through a padded, warp-private shared-memory tile, and every lane runs a second 32-point FFT.
//   in : v[r] = x[lane + 32 r]        out : v[r] = X[lane + 32 r]
// Twiddles come from a device table built once in double precision (no fast-math sincos on the data path).
#pragma once
#include <cuda_runtime.h>

namespace b200x {

constexpr int FFT_N = 1024;            // complex points per frame (real frame length 2048)
constexpr int FFT_TILE = 32 * 33;      // float2 elements of the per-warp transpose tile

// g_tw2048[j] = (cos(2 pi j / 2048), sin(2 pi j / 2048)); g_hann[n] = periodic Hann(2048);
// g_wss512[s] = sum_j hann^2[s + 512 j] (steady-state overlap-add envelope for hop 512)
__device__ float2 g_tw2048[2048];
__device__ float g_hann[2048];
__device__ float g_wss512[512];
__device__ float g_rwss512[512];     // 1 / g_wss512

__device__ __forceinline__ constexpr int brev5(int i) {
    return ((i & 1) << 4) | ((i & 2) << 2) | (i & 4) | ((i & 8) >> 2) | ((i & 16) >> 4);
}

// cos / sin of 2 pi q / 32, q = 0..15 (namespace-scope constexpr: folded into immediates after unrolling)
constexpr float kC32[16] = {1.0f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f,
                            0.70710678118654757f, 0.55557023301960229f, 0.38268343236508984f, 0.19509032201612833f,
                            0.0f, -0.19509032201612819f, -0.38268343236508973f, -0.55557023301960196f,
                            -0.70710678118654746f, -0.83146961230254535f, -0.92387953251128674f, -0.98078528040323043f};
constexpr float kS32[16] = {0.0f, 0.19509032201612825f, 0.38268343236508978f, 0.55557023301960218f,
                            0.70710678118654746f, 0.83146961230254524f, 0.92387953251128674f, 0.98078528040323043f,
                            1.0f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254546f,
                            0.70710678118654757f, 0.55557023301960218f, 0.38268343236508989f, 0.19509032201612861f};


// Per-lane trigonometry for the frame pre / post passes.  Bin k = lane + 32 r and sample pair (2m, 2m+1), m = lane + 32 r,
// both factor into a per-lane angle and a per-r angle whose cos / sin are compile-time constants after unrolling, so
//   e^(i 2 pi k / 2048)        = e^(i 2 pi lane / 2048) . e^(i 2 pi r / 64)
//   hann(n) = 1/2 - 1/2 cos(2 pi n / 2048),  cos(2 pi (2m + j) / 2048) = Re[e^(i 2 pi (2 lane + j) / 2048) . e^(i 2 pi r / 32)]
// cost 2-4 FMAs per value instead of a table load per value: the 64 loads per frame and lane (twiddle + window) were
// the top stall of the iSTFT / mel kernels (ncu: 44-54 % of samples on the long scoreboard at their first consumer).
// (constant memory: with the loops unrolled the index is an immediate, so each value is a constant-bank FMA operand)
static __constant__ float c_cos64[32] = {1.0f, 0.995184727f, 0.98078528f, 0.956940336f, 0.923879533f, 0.881921264f, 0.831469612f, 0.773010453f, 0.707106781f, 0.634393284f, 0.555570233f, 0.471396737f, 0.382683432f, 0.290284677f, 0.195090322f, 0.0980171403f, 0.0f, -0.0980171403f, -0.195090322f, -0.290284677f, -0.382683432f, -0.471396737f, -0.555570233f, -0.634393284f, -0.707106781f, -0.773010453f, -0.831469612f, -0.881921264f, -0.923879533f, -0.956940336f, -0.98078528f, -0.995184727f};
static __constant__ float c_sin64[32] = {0.0f, 0.0980171403f, 0.195090322f, 0.290284677f, 0.382683432f, 0.471396737f, 0.555570233f, 0.634393284f, 0.707106781f, 0.773010453f, 0.831469612f, 0.881921264f, 0.923879533f, 0.956940336f, 0.98078528f, 0.995184727f, 1.0f, 0.995184727f, 0.98078528f, 0.956940336f, 0.923879533f, 0.881921264f, 0.831469612f, 0.773010453f, 0.707106781f, 0.634393284f, 0.555570233f, 0.471396737f, 0.382683432f, 0.290284677f, 0.195090322f, 0.0980171403f};
static __constant__ float c_cos32[32] = {1.0f, 0.98078528f, 0.923879533f, 0.831469612f, 0.707106781f, 0.555570233f, 0.382683432f, 0.195090322f, 0.0f, -0.195090322f, -0.382683432f, -0.555570233f, -0.707106781f, -0.831469612f, -0.923879533f, -0.98078528f, -1.0f, -0.98078528f, -0.923879533f, -0.831469612f, -0.707106781f, -0.555570233f, -0.382683432f, -0.195090322f, 0.0f, 0.195090322f, 0.382683432f, 0.555570233f, 0.707106781f, 0.831469612f, 0.923879533f, 0.98078528f};
static __constant__ float c_sin32[32] = {0.0f, 0.195090322f, 0.382683432f, 0.555570233f, 0.707106781f, 0.831469612f, 0.923879533f, 0.98078528f, 1.0f, 0.98078528f, 0.923879533f, 0.831469612f, 0.707106781f, 0.555570233f, 0.382683432f, 0.195090322f, 0.0f, -0.195090322f, -0.382683432f, -0.555570233f, -0.707106781f, -0.831469612f, -0.923879533f, -0.98078528f, -1.0f, -0.98078528f, -0.923879533f, -0.831469612f, -0.707106781f, -0.555570233f, -0.382683432f, -0.195090322f};

struct LaneTrig {
    float2 tw;      // e^(i 2 pi lane / 2048)
    float2 h0, h1;  // 1/2 e^(i 2 pi (2 lane) / 2048), 1/2 e^(i 2 pi (2 lane + 1) / 2048)
};
__device__ __forceinline__ LaneTrig lane_trig(int lane) {
    LaneTrig t;
    t.tw = g_tw2048[lane];
    const float2 a = g_tw2048[2 * lane], b = g_tw2048[2 * lane + 1];
    t.h0 = make_float2(0.5f * a.x, 0.5f * a.y);
    t.h1 = make_float2(0.5f * b.x, 0.5f * b.y);
    return t;
}
// (cos, sin)(2 pi (lane + 32 r) / 2048)
__device__ __forceinline__ float2 twiddle2048(const LaneTrig& t, int r) {
    const float cr = c_cos64[r], sr = c_sin64[r];
    return make_float2(fmaf(t.tw.x, cr, -t.tw.y * sr), fmaf(t.tw.y, cr, t.tw.x * sr));
}
// periodic Hann(2048) at samples 2m and 2m + 1, m = lane + 32 r
__device__ __forceinline__ float2 hann_pair(const LaneTrig& t, int r) {
    const float cr = c_cos32[r], sr = c_sin32[r];
    return make_float2(fmaf(t.h0.y, sr, fmaf(-t.h0.x, cr, 0.5f)), fmaf(t.h1.y, sr, fmaf(-t.h1.x, cr, 0.5f)));
}

template <bool INV, int K, int J, int M>
__device__ __forceinline__ void fft32_bfly(float2 (&t)[32]) {
    constexpr int H = M / 2;
    constexpr int Q = J * (32 / M);
    const float2 x = t[K + J + H];
    float2 wx;
    if constexpr (Q == 0) {
        wx = x;
    } else if constexpr (Q == 8) {
        wx = INV ? make_float2(-x.y, x.x) : make_float2(x.y, -x.x);
    } else {
        constexpr float wr = kC32[Q];
        constexpr float wi = INV ? kS32[Q] : -kS32[Q];
        wx = make_float2(x.x * wr - x.y * wi, x.x * wi + x.y * wr);
    }
    const float2 u = t[K + J];
    t[K + J] = make_float2(u.x + wx.x, u.y + wx.y);
    t[K + J + H] = make_float2(u.x - wx.x, u.y - wx.y);
}

template <bool INV, int M, int I>
__device__ __forceinline__ void fft32_stage(float2 (&t)[32]) {
    // butterfly I of 16 in the stage with span M: group K = (I / (M/2)) * M, offset J = I % (M/2)
    if constexpr (I < 16) {
        fft32_bfly<INV, (I / (M / 2)) * M, I % (M / 2), M>(t);
        fft32_stage<INV, M, I + 1>(t);
    }
}

template <bool INV>
__device__ __forceinline__ void fft32(float2 (&v)[32]) {
    float2 t[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) t[i] = v[brev5(i)];
    fft32_stage<INV, 2, 0>(t);
    fft32_stage<INV, 4, 0>(t);
    fft32_stage<INV, 8, 0>(t);
    fft32_stage<INV, 16, 0>(t);
    fft32_stage<INV, 32, 0>(t);
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = t[i];
}

constexpr int FFT_TWIDDLE = 32 * 32;   // float2 elements of the CTA-wide inter-stage twiddle table

// tw[k2 * 32 + lane] = (cos, sin)(2 pi lane k2 / 1024): the inter-stage twiddles depend only on (lane, k2), so one
// 8 KB shared-memory table per CTA serves every frame with conflict-free 8-byte loads.  (Gathering them from the global
// table costs up to 32 L1 wavefronts per load - lanes stride 16 k2 bytes - and was the top stall of the FFT kernels.)
__device__ __forceinline__ void fft_fill_twiddles(float2* tw) {
    for (int i = threadIdx.x; i < FFT_TWIDDLE; i += blockDim.x) {
        const int k2 = i >> 5, n1 = i & 31;
        tw[i] = g_tw2048[(2 * n1 * k2) & 2047];
    }
    __syncthreads();
}

template <bool INV>
__device__ __forceinline__ void fft1024_warp(float2 (&v)[32], float2* tile, const float2* tw, int lane) {
    // ONE copy of the unrolled 32-point FFT in the instruction stream, executed twice: the fully inlined form of the frame
    // kernels was ~50 KB of code, past the instruction cache (ncu: 25 % of the mel kernel's samples were "no instruction")
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        fft32<INV>(v);                                 // pass 0: over n2 (lane = n1): v[k2];  pass 1: over n1 (lane = k2): v[k1] = X[32 k1 + k2]
        if (pass == 0) {
#pragma unroll
            for (int k2 = 1; k2 < 32; ++k2) {          // times W_1024^(n1 k2)
                const float2 w = tw[k2 * 32 + lane];
                const float wi = INV ? w.y : -w.y;
                const float2 x = v[k2];
                v[k2] = make_float2(x.x * w.x - x.y * wi, x.x * wi + x.y * w.x);
            }
            __syncwarp();
#pragma unroll
            for (int k2 = 0; k2 < 32; ++k2) tile[k2 * 33 + lane] = v[k2];
            __syncwarp();
#pragma unroll
            for (int n1 = 0; n1 < 32; ++n1) v[n1] = tile[lane * 33 + n1];
        }
    }
}

}  // namespace b200x
