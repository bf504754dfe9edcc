// Fused multi-head self-attention for the SpecTTTra encoder on sm_100a (head_dim 64, bf16 in, fp32 accumulate).
//
// One CTA per (128-query tile, head, perturbed copy); two CTAs are co-resident per SM so one CTA's softmax
// overlaps the other's MMAs.  Per CTA:
//   warp 0 : TMA producer  - Q tile once, K / V tiles through a 2-stage ring (3-D tensor map over [copy][token][1152])
//   warp 1 : TMEM allocator + tcgen05.mma issuer:  S = Q.K^T (SS form)  then  O += P.V (P read from TMEM, V MN-major)
//   warps 2-5 : softmax - one query row per thread (TMEM lane == row, so row max / sum need no shuffles); exp2 with the
//               1/sqrt(d) scale folded in, lazy (thresholded) rescaling of O, P written back over S in TMEM as bf16.
// The qkv buffer is the QKV GEMM output [copies * tokens, 3 * heads * 64] = [q | k | v] (timm reshape order).
#include "common.h"
#include "ptx.cuh"

namespace b200x {

constexpr int ATT_THREADS = 192;
constexpr int ATT_TILE = 128;
constexpr int ATT_HD = 64;
constexpr int ATT_TILE_BYTES = ATT_TILE * ATT_HD * 2;     // 16 KB
constexpr int ATT_KV_STAGES = 2;
constexpr int ATT_SMEM = ATT_TILE_BYTES * (1 + 2 * ATT_KV_STAGES) + 256 + 1024;
constexpr uint32_t ATT_TMEM_COLS = 256;                  // S: [0,128)  O: [128,192)
constexpr uint32_t ATT_O_COL = 128;

struct AttnParams {
    int tokens;        // tokens per copy (multiple of 16)
    int heads;
    __nv_bfloat16* out;   // [copies * tokens, heads * 64]
    float scale_log2;     // (1/sqrt(64)) * log2(e)
};

__global__ void __launch_bounds__(ATT_THREADS, 2)
attention_kernel(const __grid_constant__ CUtensorMap tmQKV, AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sK = smem + ATT_TILE_BYTES;
    uint8_t* sV = sK + ATT_KV_STAGES * ATT_TILE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + ATT_KV_STAGES * ATT_TILE_BYTES);
    uint64_t* q_full = bars;
    uint64_t* kv_full = bars + 1;
    uint64_t* kv_empty = kv_full + ATT_KV_STAGES;
    uint64_t* s_full = kv_empty + ATT_KV_STAGES;
    uint64_t* p_ready = s_full + 1;
    uint64_t* o_done = p_ready + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int q_tile = blockIdx.x, head = blockIdx.y, copy = blockIdx.z;
    const int q0 = q_tile * ATT_TILE;
    const int nkv = (p.tokens + ATT_TILE - 1) / ATT_TILE;
    const int hidden = p.heads * ATT_HD;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQKV);
        mbar_init(q_full, 1);
        for (int s = 0; s < ATT_KV_STAGES; ++s) {
            mbar_init(&kv_full[s], 1);
            mbar_init(&kv_empty[s], 1);
        }
        mbar_init(s_full, 1);
        mbar_init(p_ready, 4);
        mbar_init(o_done, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<ATT_TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(q_full, ATT_TILE_BYTES);
            tma_load_3d(sQ, &tmQKV, q_full, head * ATT_HD, q0, copy);
            for (int j = 0; j < nkv; ++j) {
                const int st = j % ATT_KV_STAGES;
                const uint32_t ph = (j / ATT_KV_STAGES) & 1;
                mbar_wait(&kv_empty[st], ph ^ 1);
                mbar_expect_tx(&kv_full[st], 2 * ATT_TILE_BYTES);
                tma_load_3d(sK + st * ATT_TILE_BYTES, &tmQKV, &kv_full[st], hidden + head * ATT_HD, j * ATT_TILE, copy);
                tma_load_3d(sV + st * ATT_TILE_BYTES, &tmQKV, &kv_full[st], 2 * hidden + head * ATT_HD, j * ATT_TILE, copy);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc_pv = make_idesc_bf16(ATT_TILE, ATT_HD, true);
            const uint32_t q_addr = smem_u32(sQ);
            const uint32_t tS = tmem_base, tO = tmem_base + ATT_O_COL;
            mbar_wait(q_full, 0);
            for (int j = 0; j < nkv; ++j) {
                const int st = j % ATT_KV_STAGES;
                const uint32_t ph = (j / ATT_KV_STAGES) & 1;
                const int nk = min(ATT_TILE, p.tokens - j * ATT_TILE);
                const uint32_t idesc_s = make_idesc_bf16(ATT_TILE, nk, false);
                mbar_wait(&kv_full[st], ph);
                tc_fence_after();
                const uint32_t k_addr = smem_u32(sK + st * ATT_TILE_BYTES);
                const uint32_t v_addr = smem_u32(sV + st * ATT_TILE_BYTES);
#pragma unroll
                for (int k = 0; k < ATT_HD / 16; ++k)
                    umma_ss(tS, make_smem_desc_sw128(q_addr + k * 32, 16, 1024),
                            make_smem_desc_sw128(k_addr + k * 32, 16, 1024), idesc_s, k != 0 ? 1u : 0u);
                umma_commit(s_full);
                mbar_wait(p_ready, j & 1);
                tc_fence_after();
                for (int ks = 0; ks < nk / 16; ++ks)
                    umma_ts(tO, tS + ks * 8, make_smem_desc_sw128(v_addr + ks * 2048, 16384, 1024), idesc_pv,
                            (j | ks) != 0 ? 1u : 0u);
                umma_commit(&kv_empty[st]);
                umma_commit(o_done);
            }
        }
    } else {
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
        const float c = p.scale_log2;
        float m_ref = -INFINITY, l_sum = 0.f;
        for (int j = 0; j < nkv; ++j) {
            const int nk = min(ATT_TILE, p.tokens - j * ATT_TILE);
            mbar_wait(s_full, j & 1);
            tc_fence_after();
            float s[ATT_TILE];
#pragma unroll
            for (int cidx = 0; cidx < ATT_TILE; cidx += 16)
                if (cidx < nk) tmem_ld16(t_lane + cidx, reinterpret_cast<uint32_t*>(&s[cidx]));
            tmem_wait_ld();
            float mt = -INFINITY;
#pragma unroll
            for (int i = 0; i < ATT_TILE; ++i)
                if (i < nk) mt = fmaxf(mt, s[i]);
            // lazy rescale: keep the old reference max unless the row max grew by more than 2^8
            const bool need = (j == 0) || ((mt - m_ref) * c > 8.0f);
            const float m_new = need ? fmaxf(mt, m_ref) : m_ref;
            if (j > 0 && __any_sync(0xffffffffu, need)) {
                // S_j complete implies PV_{j-1} complete (same in-order tensor pipe), so O is quiescent here
                const float sc = need ? ex2_approx((m_ref - m_new) * c) : 1.0f;
                l_sum *= sc;
#pragma unroll
                for (int cidx = 0; cidx < ATT_HD; cidx += 16) {
                    uint32_t o[16];
                    tmem_ld16(t_lane + ATT_O_COL + cidx, o);
                    tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * sc);
                    tmem_st16(t_lane + ATT_O_COL + cidx, o);
                }
            }
            m_ref = m_new;
            const float mc = m_ref * c;
            float acc = 0.f;
#pragma unroll
            for (int cidx = 0; cidx < ATT_TILE; cidx += 32) {
                if (cidx < nk) {
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int a = cidx + 2 * i;
                        float e0 = 0.f, e1 = 0.f;
                        if (a < nk) { e0 = ex2_approx(fmaf(s[a], c, -mc)); e1 = ex2_approx(fmaf(s[a + 1], c, -mc)); }
                        acc += e0 + e1;
                        pk[i] = pack_bf16(e0, e1);
                    }
                    tmem_st16(t_lane + cidx / 2, pk);
                }
            }
            l_sum += acc;
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_ready);
        }
        mbar_wait(o_done, (nkv - 1) & 1);
        tc_fence_after();
        const int q = q0 + row;
        const float inv = 1.0f / l_sum;
        uint4 packed[8];
#pragma unroll
        for (int cidx = 0; cidx < ATT_HD; cidx += 16) {
            uint32_t o[16];
            tmem_ld16(t_lane + ATT_O_COL + cidx, o);
            tmem_wait_ld();
            uint32_t w[8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
                w[i] = pack_bf16(__uint_as_float(o[2 * i]) * inv, __uint_as_float(o[2 * i + 1]) * inv);
            packed[cidx / 8] = make_uint4(w[0], w[1], w[2], w[3]);
            packed[cidx / 8 + 1] = make_uint4(w[4], w[5], w[6], w[7]);
        }
        if (q < p.tokens) {
            uint4* dst = reinterpret_cast<uint4*>(p.out + (static_cast<long long>(copy) * p.tokens + q) * hidden + head * ATT_HD);
#pragma unroll
            for (int i = 0; i < 8; ++i) dst[i] = packed[i];
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<ATT_TMEM_COLS>(tmem_base);
}

}  // namespace b200x

using namespace b200x;

extern "C" int b200x_attention(const void* d_qkv, void* d_out, int copies, int tokens, int heads, int head_dim,
                               void* stream) {
    B200X_REQUIRE(head_dim == ATT_HD, "attention: head_dim %d unsupported (kernel is specialised for 64)", head_dim);
    B200X_REQUIRE(copies > 0 && tokens > 0 && heads > 0, "attention: empty problem");
    B200X_REQUIRE(tokens % 16 == 0, "attention: tokens=%d must be a multiple of 16", tokens);
    static bool configured = false;
    if (!configured) {
        B200X_CUDA_TRY(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
        configured = true;
    }
    const int width = 3 * heads * ATT_HD;
    CUtensorMap tm;
    const uint64_t dims[3] = {static_cast<uint64_t>(width), static_cast<uint64_t>(tokens), static_cast<uint64_t>(copies)};
    const uint64_t strides[2] = {static_cast<uint64_t>(width) * 2, static_cast<uint64_t>(width) * 2 * tokens};
    const uint32_t box[3] = {ATT_HD, ATT_TILE, 1};
    B200X_TRY(make_tmap_bf16(&tm, d_qkv, 3, dims, strides, box));
    AttnParams p{tokens, heads, reinterpret_cast<__nv_bfloat16*>(d_out), 0.125f * 1.4426950408889634f};
    dim3 grid(ceil_div(tokens, ATT_TILE), heads, copies);
    attention_kernel<<<grid, ATT_THREADS, ATT_SMEM, static_cast<cudaStream_t>(stream)>>>(tm, p);
    B200X_CUDA_TRY(cudaGetLastError());
    return B200X_OK;
}
