// Fused multi-head self-attention for the SpecTTTra encoder on sm_100a (head_dim 64, bf16 in, fp32 accumulate).
//
// One CTA per (256-query block, head, perturbed copy), one CTA per SM.  The block is two 128-row query tiles (A, B) that
// share every K / V tile brought in by TMA (halves the L2 -> shared-memory traffic per query row) and ping-pong on the
// tensor pipe: while the softmax warps of tile A work on S_A, the MMAs of tile B run, and vice versa.
//   warp 0     : TMA producer - Q_A, Q_B once; K / V tiles through a 4-stage mbarrier ring
//                (3-D tensor map over [copy][token][3 * heads * 64], 128-byte swizzle, OOB rows zero-filled)
//   warp 1, 10 : one tcgen05.mma issuer thread per query tile (warp 1 also owns the TMEM allocation):
//                S_x = Q_x K^T (SS form, fp32 in TMEM),  O_x += P_x V (P_x read from TMEM, V MN-major in shared memory);
//                two issuers keep the tiles' dependency chains (softmax -> PV -> next S) independent of each other
//   warps 2-5  : softmax of tile A, warps 6-9: softmax of tile B.  One query row per thread (TMEM lane == row, so row
//                max / sum need no shuffles); exp2 with 1/sqrt(d) folded in; the running reference max is only replaced
//                (and O rescaled) when a tile's row max exceeds it by more than 2^8, otherwise scores stream through
//                TMEM -> exp2 -> bf16 P -> TMEM in a single pass.
// TMEM columns: S_A [0,128) S_B [128,256) O_A [256,320) O_B [320,384) P_A [384,448) P_B [448,512).
// The qkv buffer is the QKV GEMM output [copies * tokens, 3 * heads * 64] = [q | k | v] (timm reshape order).
#include "common.h"
#include "ptx.cuh"

namespace b200x {

constexpr int ATT_THREADS = 352;       // warp 0 TMA, 1 MMA(A), 2-5 softmax(A), 6-9 softmax(B), 10 MMA(B)
constexpr int ATT_TILE = 128;
constexpr int ATT_HD = 64;
constexpr int ATT_TILE_BYTES = ATT_TILE * ATT_HD * 2;     // 16 KB
constexpr int ATT_KV_STAGES = 4;
constexpr int ATT_SMEM = ATT_TILE_BYTES * (2 + 2 * ATT_KV_STAGES) + 256 + 1024;
constexpr uint32_t ATT_TMEM_COLS = 512;
constexpr uint32_t ATT_S_COL = 0, ATT_O_COL = 256, ATT_P_COL = 384;
constexpr float ATT_RESCALE_LOG2 = 8.0f;

struct AttnParams {
    int tokens;        // tokens per copy (multiple of 16)
    int heads;
    __nv_bfloat16* out;   // [copies * tokens, heads * 64]
    float scale_log2;     // (1/sqrt(64)) * log2(e)
};

// 32 (or, for a tail chunk, 16) scores -> p = 2^(s*c - m_ref*c) -> packed bf16 in TMEM; accumulates the row sum / tile max.
// The fp32 -> bf16 conversion is a byte permute that keeps the high halves (truncation) instead of F2FP: the conversion
// unit shares the MUFU pipe, which is what bounds this kernel.  The row sum is taken over the TRUNCATED values, so the
// softmax weights stay exactly normalised and carry the same error variance as round-to-nearest, without bias.
__device__ __forceinline__ void softmax_chunk(const uint32_t (&r)[32], bool wide, uint32_t tP_col, float c, float mc,
                                              float& acc, float& mt, int dbg = 0) {
    uint32_t pk[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        uint32_t b0 = 0u, b1 = 0u;
        if (wide || i < 8) {
            const float s0 = __uint_as_float(r[2 * i]), s1 = __uint_as_float(r[2 * i + 1]);
            mt = fmax3(mt, s0, s1);
            if (dbg == 1) {
                b0 = __float_as_uint(fmaf(s0, c, -mc)) & 0xFFFF0000u;
                b1 = __float_as_uint(fmaf(s1, c, -mc)) & 0xFFFF0000u;
            } else {
                b0 = __float_as_uint(ex2_approx(fmaf(s0, c, -mc))) & 0xFFFF0000u;
                b1 = __float_as_uint(ex2_approx(fmaf(s1, c, -mc))) & 0xFFFF0000u;
            }
        }
        acc += __uint_as_float(b0) + __uint_as_float(b1);
        pk[i] = __byte_perm(b0, b1, 0x7632);      // low half = bf16(s0), high half = bf16(s1)
    }
    if (dbg != 3) tmem_st16(tP_col, pk);       // a 16-column tail chunk stores 8 meaningful + 8 zero words (never read)
}

// One streaming pass over a score tile.  The TMEM load of chunk i+1 is in flight while chunk i is exponentiated
// (two register buffers; a buffer is only read after the tcgen05.wait::ld that follows its load).
__device__ __forceinline__ void softmax_pass(uint32_t tS, uint32_t tP, int nk, float c, float mc, float& acc, float& mt, int dbg = 0) {
    acc = 0.f;
    mt = -INFINITY;
    if (dbg == 4) { acc = 1.f; mt = mc / c; return; }
    uint32_t ra[32], rb[32];
    if (dbg == 2) {
#pragma unroll
        for (int i = 0; i < 32; ++i) ra[i] = rb[i] = __float_as_uint(mc / c - 0.001f * i);
    }
    if (dbg != 2) { if (nk >= 32) tmem_ld32(tS, ra); else tmem_ld16(tS, ra); }
#pragma unroll
    for (int ci = 0; ci < ATT_TILE / 32; ++ci) {
        const int col = ci * 32;
        if (col < nk) {
            tmem_wait_ld();
            const int nxt = col + 32;
            if (nxt < nk && dbg != 2) {
                if ((ci & 1) == 0) { if (nxt + 32 <= nk) tmem_ld32(tS + nxt, rb); else tmem_ld16(tS + nxt, rb); }
                else               { if (nxt + 32 <= nk) tmem_ld32(tS + nxt, ra); else tmem_ld16(tS + nxt, ra); }
            }
            if ((ci & 1) == 0) softmax_chunk(ra, col + 32 <= nk, tP + col / 2, c, mc, acc, mt, dbg);
            else               softmax_chunk(rb, col + 32 <= nk, tP + col / 2, c, mc, acc, mt, dbg);
        }
    }
}

template <int DBG>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_kernel(const __grid_constant__ CUtensorMap tmQKV, AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;                                   // two tiles
    uint8_t* sK = smem + 2 * ATT_TILE_BYTES;
    uint8_t* sV = sK + ATT_KV_STAGES * ATT_TILE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + ATT_KV_STAGES * ATT_TILE_BYTES);
    uint64_t* q_full = bars;
    uint64_t* kv_full = bars + 1;
    uint64_t* kv_empty = kv_full + ATT_KV_STAGES;
    uint64_t* s_full = kv_empty + ATT_KV_STAGES;          // [2]
    uint64_t* p_ready = s_full + 2;                       // [2]
    uint64_t* o_done = p_ready + 2;                       // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int head = blockIdx.y, copy = blockIdx.z;
    const int q0 = blockIdx.x * 2 * ATT_TILE;
    const int nq = (q0 + ATT_TILE < p.tokens) ? 2 : 1;    // query tiles of this block that hold valid rows
    const int nkv = (p.tokens + ATT_TILE - 1) / ATT_TILE;
    const int hidden = p.heads * ATT_HD;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQKV);
        mbar_init(q_full, 1);
        for (int s = 0; s < ATT_KV_STAGES; ++s) {
            mbar_init(&kv_full[s], 1);
            mbar_init(&kv_empty[s], nq);
        }
        for (int x = 0; x < 2; ++x) {
            mbar_init(&s_full[x], 1);
            mbar_init(&p_ready[x], 4);
            mbar_init(&o_done[x], 1);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<ATT_TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(q_full, nq * ATT_TILE_BYTES);
            tma_load_3d(sQ, &tmQKV, q_full, head * ATT_HD, q0, copy);
            if (nq == 2) tma_load_3d(sQ + ATT_TILE_BYTES, &tmQKV, q_full, head * ATT_HD, q0 + ATT_TILE, copy);
            for (int j = 0; j < nkv; ++j) {
                const int st = j % ATT_KV_STAGES;
                const uint32_t ph = (j / ATT_KV_STAGES) & 1;
                mbar_wait(&kv_empty[st], ph ^ 1);
                mbar_expect_tx(&kv_full[st], 2 * ATT_TILE_BYTES);
                tma_load_3d(sK + st * ATT_TILE_BYTES, &tmQKV, &kv_full[st], hidden + head * ATT_HD, j * ATT_TILE, copy);
                tma_load_3d(sV + st * ATT_TILE_BYTES, &tmQKV, &kv_full[st], 2 * hidden + head * ATT_HD, j * ATT_TILE, copy);
            }
        }
    } else if (warp == 1 || warp == 10) {
        const int x = (warp == 1) ? 0 : 1;                // query tile served by this issuer
        if (lane == 0 && x < nq) {
            constexpr uint32_t idesc_pv = make_idesc_bf16(ATT_TILE, ATT_HD, true);
            const uint32_t tS = tmem_base + ATT_S_COL + x * ATT_TILE;
            const uint32_t tO = tmem_base + ATT_O_COL + x * ATT_HD;
            const uint32_t tP = tmem_base + ATT_P_COL + x * ATT_HD;
            const uint64_t q_desc = make_smem_desc_sw128(smem_u32(sQ) + x * ATT_TILE_BYTES, 16, 1024);
            const uint64_t k_desc0 = make_smem_desc_sw128(smem_u32(sK), 16, 1024);
            const uint64_t v_desc0 = make_smem_desc_sw128(smem_u32(sV), 16384, 1024);
            auto issue_s = [&](int j) {                   // S_x = Q_x K_j^T
                const int nk = min(ATT_TILE, p.tokens - j * ATT_TILE);
                const uint32_t idesc_s = make_idesc_bf16(ATT_TILE, nk, false);
                const uint64_t kd = k_desc0 + static_cast<uint64_t>((j % ATT_KV_STAGES) * (ATT_TILE_BYTES >> 4));
                if (!(DBG & 16))
#pragma unroll
                for (int k = 0; k < ATT_HD / 16; ++k) umma_ss(tS, q_desc + 2 * k, kd + 2 * k, idesc_s, k != 0 ? 1u : 0u);
                umma_commit(&s_full[x]);
            };
            auto issue_pv = [&](int j) {                  // O_x += P_x V_j
                const int nk = min(ATT_TILE, p.tokens - j * ATT_TILE);
                const uint64_t vd = v_desc0 + static_cast<uint64_t>((j % ATT_KV_STAGES) * (ATT_TILE_BYTES >> 4));
                if (DBG & 8) return;
                if (nk == ATT_TILE) {
#pragma unroll
                    for (int ks = 0; ks < ATT_TILE / 16; ++ks) umma_ts(tO, tP + ks * 8, vd + 128 * ks, idesc_pv, (j | ks) != 0 ? 1u : 0u);
                } else {
                    for (int ks = 0; ks < nk / 16; ++ks) umma_ts(tO, tP + ks * 8, vd + 128 * ks, idesc_pv, (j | ks) != 0 ? 1u : 0u);
                }
            };
            mbar_wait(q_full, 0);
            mbar_wait(&kv_full[0], 0);
            tc_fence_after();
            issue_s(0);
            for (int j = 0; j < nkv; ++j) {
                mbar_wait(&p_ready[x], j & 1);
                tc_fence_after();
                issue_pv(j);
                umma_commit(&kv_empty[j % ATT_KV_STAGES]);              // this tile is done with K_j / V_j
                if (j + 1 < nkv) {
                    mbar_wait(&kv_full[(j + 1) % ATT_KV_STAGES], ((j + 1) / ATT_KV_STAGES) & 1);
                    tc_fence_after();
                    issue_s(j + 1);                                      // its commit also covers the PV just issued
                } else {
                    umma_commit(&o_done[x]);
                }
            }
        }
    } else {
        const int x = (warp - 2) >> 2;                    // query tile of this softmax warp group (warps 2-5: A, 6-9: B)
        if (x < nq) {
            const int quarter = warp & 3;
            const int row = quarter * 32 + lane;
            const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
            const uint32_t tS = t_lane + ATT_S_COL + x * ATT_TILE;
            const uint32_t tO = t_lane + ATT_O_COL + x * ATT_HD;
            const uint32_t tP = t_lane + ATT_P_COL + x * ATT_HD;
            const float c = p.scale_log2;
            float m_ref = -INFINITY, l_sum = 0.f;
            for (int j = 0; j < nkv; ++j) {
                const int nk = min(ATT_TILE, p.tokens - j * ATT_TILE);
                mbar_wait(&s_full[x], j & 1);
                tc_fence_after();
                if (j == 0 && (DBG & 7) != 4 && (DBG & 7) != 2) { // first tile: exact row max first
                    float mt = -INFINITY;
#pragma unroll 1
                    for (int col = 0; col < nk; col += 16) {
                        uint32_t r[16];
                        tmem_ld16(tS + col, r);
                        tmem_wait_ld();
#pragma unroll
                        for (int i = 0; i < 16; ++i) mt = fmaxf(mt, __uint_as_float(r[i]));
                    }
                    m_ref = mt;
                }
                if (j == 0 && ((DBG & 7) == 4 || (DBG & 7) == 2)) m_ref = 0.f;
                float acc, mt;
                softmax_pass(tS, tP, nk, c, m_ref * c, acc, mt, DBG & 7);
                const bool need = (mt - m_ref) * c > ATT_RESCALE_LOG2;
                if (__any_sync(0xffffffffu, need)) {
                    // rare: adopt the larger max, rescale the running sum and O (S_j complete implies PV_{j-1} complete
                    // on the in-order tensor pipe, so O is quiescent), then redo the tile against the new reference.
                    const float m_new = fmaxf(m_ref, mt);
                    const float sc = ex2_approx((m_ref - m_new) * c);
                    l_sum *= sc;
                    if (j > 0) {
#pragma unroll
                        for (int cidx = 0; cidx < ATT_HD; cidx += 16) {
                            uint32_t o[16];
                            tmem_ld16(tO + cidx, o);
                            tmem_wait_ld();
#pragma unroll
                            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * sc);
                            tmem_st16(tO + cidx, o);
                        }
                    }
                    m_ref = m_new;
                    softmax_pass(tS, tP, nk, c, m_ref * c, acc, mt);
                }
                l_sum += acc;
                tmem_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_ready[x]);
            }
            mbar_wait(&o_done[x], 0);
            tc_fence_after();
            const int q = q0 + x * ATT_TILE + row;
            const float inv = 1.0f / l_sum;
            uint4 packed[8];
#pragma unroll
            for (int cidx = 0; cidx < ATT_HD; cidx += 16) {
                uint32_t o[16];
                tmem_ld16(tO + cidx, o);
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
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<ATT_TMEM_COLS>(tmem_base);
}

}  // namespace b200x

using namespace b200x;

static int g_attn_dbg = 0;

template <int DBG>
static int launch_attention(const CUtensorMap& tm, const AttnParams& p, dim3 grid, cudaStream_t s) {
    static bool configured = false;
    if (!configured) {
        B200X_CUDA_TRY(cudaFuncSetAttribute(attention_kernel<DBG>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
        configured = true;
    }
    attention_kernel<DBG><<<grid, ATT_THREADS, ATT_SMEM, s>>>(tm, p);
    B200X_CUDA_TRY(cudaGetLastError());
    return B200X_OK;
}

// diagnostic only (not part of the public header): select a stripped-down variant of the kernel for bottleneck analysis
extern "C" void b200x_debug_attention_variant(int v) { g_attn_dbg = v; }

extern "C" int b200x_attention(const void* d_qkv, void* d_out, int copies, int tokens, int heads, int head_dim,
                               void* stream) {
    B200X_REQUIRE(head_dim == ATT_HD, "attention: head_dim %d unsupported (kernel is specialised for 64)", head_dim);
    B200X_REQUIRE(copies > 0 && tokens > 0 && heads > 0, "attention: empty problem");
    B200X_REQUIRE(tokens % 16 == 0, "attention: tokens=%d must be a multiple of 16", tokens);
    const int width = 3 * heads * ATT_HD;
    CUtensorMap tm;
    const uint64_t dims[3] = {static_cast<uint64_t>(width), static_cast<uint64_t>(tokens), static_cast<uint64_t>(copies)};
    const uint64_t strides[2] = {static_cast<uint64_t>(width) * 2, static_cast<uint64_t>(width) * 2 * tokens};
    const uint32_t box[3] = {ATT_HD, ATT_TILE, 1};
    B200X_TRY(make_tmap_bf16(&tm, d_qkv, 3, dims, strides, box));
    AttnParams p{tokens, heads, reinterpret_cast<__nv_bfloat16*>(d_out), 0.125f * 1.4426950408889634f};
    dim3 grid(ceil_div(tokens, 2 * ATT_TILE), heads, copies);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    switch (g_attn_dbg) {
        case 0: return launch_attention<0>(tm, p, grid, s);
        case 4: return launch_attention<4>(tm, p, grid, s);
        case 12: return launch_attention<12>(tm, p, grid, s);
        case 20: return launch_attention<20>(tm, p, grid, s);
        case 28: return launch_attention<28>(tm, p, grid, s);
        case 1: return launch_attention<1>(tm, p, grid, s);
        case 2: return launch_attention<2>(tm, p, grid, s);
        case 3: return launch_attention<3>(tm, p, grid, s);
        default: return set_error(B200X_ERR_INVALID, "attention: unknown diagnostic variant %d", g_attn_dbg);
    }
}
