// Fused multi-head self-attention for the SpecTTTra encoder on sm_100a (head_dim 64, bf16 in, fp32 accumulate).
//
// Replaces F.scaled_dot_product_attention inside the third-party `sonics` encoder block that the reference reaches through
// LocalSonnics.predict (src/sonics_api.py:259-271; architecture restated in SURVEY.md 3d).
//
// One CTA per (128-query tile, head, perturbed copy) and TWO CTAs per SM (112 KB of shared memory, 256 TMEM columns and
// 168 registers x 192 threads each): the co-resident CTAs run free of each other, so one CTA's load / row-max / barrier
// phases and its prologue / epilogue fall into the other's exponential phase.
//   warps 0-3 : softmax.  One query row per thread (TMEM lane == row, so row max / sum need no shuffles); exp2 with
//               1/sqrt(d) folded in; the running reference max is only replaced (and O rescaled) when a tile's row max
//               exceeds it by more than 2^8, otherwise scores stream TMEM -> exp2 -> bf16 P -> TMEM in a single pass.
//               A quarter of the exponentials are evaluated on the FMA pipe (exp2_poly2): with head_dim 64 the kernel needs
//               128 ex2 per 2 x 128 x 64 MACs per row and tile, i.e. twice as many MUFU cycles as tensor cycles.
//   warp 4    : TMA producer - Q once; K / V tiles through a 3-stage mbarrier ring
//               (3-D tensor map over [copy][token][3 * heads * 64], 128-byte swizzle, OOB rows zero-filled)
//   warp 5    : TMEM allocator + single-thread tcgen05.mma issuer:  S = Q K^T (SS form, fp32 in TMEM), handed back as soon
//               as the softmax warps hold the row in registers so that S(j+1) is computed during the exponentials of tile j;
//               O += P V (P read from TMEM, V MN-major in shared memory).
//               The producer / issuer roles sit on the HIGHEST warp ids: the SM's warp arbiter prefers higher ids, and an
//               issuer that loses its issue slots to the softmax warps delays every tcgen05.mma by hundreds of cycles.
// TMEM columns: S [0,128)  O [128,192)  P [192,256).
// The qkv buffer is the QKV GEMM output [copies * tokens, 3 * heads * 64] = [q | k | v] (timm reshape order).
//
// Variants that were built, measured and rejected (two tiles per CTA, split rows, 64-key tiles, software-pipelined loads,
// a persistent ping-pong kernel with one softmax warp set alternating between two tiles, 64-key tiles with three CTAs per SM
// and P beside S) live in tools/variants/ with their
// numbers in DESIGN.md 3.2; none of them is part of this library.
#include "common.h"
#include "ptx.cuh"

namespace b200x {

constexpr int ATT_TILE = 128;
constexpr int ATT_HD = 64;
constexpr int ATT_TILE_BYTES = ATT_TILE * ATT_HD * 2;     // 16 KB
constexpr int ATT_STAGES = 3;
constexpr int ATT_THREADS = 192;
constexpr int ATT_W_TMA = 4, ATT_W_MMA = 5;
constexpr int ATT_SMEM = ATT_TILE_BYTES * (1 + 2 * ATT_STAGES) + 256;          // dynamic shared memory is 1024-byte aligned
constexpr uint32_t ATT_TMEM_COLS = 256, ATT_S_COL = 0, ATT_O_COL = 128, ATT_P_COL = 192;
constexpr float ATT_RESCALE_LOG2 = 8.0f;

struct AttnParams {
    int tokens;           // tokens per copy (multiple of 16)
    int heads;
    __nv_bfloat16* out;   // [copies * tokens, heads * 64]
    float scale_log2;     // (1/sqrt(64)) * log2(e)
    float zero;           // always 0.0f: an operand ptxas cannot fold (see exp_chunk)
    int reverse;          // walk (copy, head) from the end: L2 reuse of the QKV GEMM's last output
};

// 32 scores (registers r[0..32)) -> p = 2^(s*c - m_ref*c) -> bf16 pairs (round to nearest) -> 16 TMEM columns of P.
// Packed fp32x2 math keeps the issue cost at 2.5 slots per score: 1/2 FFMA2, MUFU, 1/2 F2FP, 1/2 FADD2; every fourth pair
// takes the FMA-pipe polynomial instead of the MUFU.
// Scheduling: with the whole row in registers ptxas would hoist all 64 FFMA2 and then emit the MUFUs in one long run,
// which blocks the (in-order) warp on the MUFU queue while its other work waits.  The two 16-score halves of a chunk
// therefore take their addend from `link` = fma(sum so far, 0, -m*c): a true data dependence on older results
// (value unchanged) that keeps at most two halves in flight, so MUFU runs stay short and interleave with FMA work.
__device__ __forceinline__ void exp_chunk(const uint32_t* r, uint32_t (&pk)[16], uint64_t c2, uint64_t nmc2, uint64_t zero2,
                                          uint64_t& acc_a, uint64_t& acc_b) {
    const uint64_t link_a = ffma2(acc_a, zero2, nmc2);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float x0, x1, p0, p1;
        unpack_f32x2(ffma2(pack_f32x2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), c2, link_a), x0, x1);
        if ((i & 3) == 3) exp2_poly2(x0, x1, p0, p1);
        else { p0 = ex2_approx(x0); p1 = ex2_approx(x1); }
        acc_a = fadd2(acc_a, pack_f32x2(p0, p1));
        pk[i] = pack_bf16(p0, p1);
    }
    const uint64_t link_b = ffma2(acc_b, zero2, nmc2);
#pragma unroll
    for (int i = 8; i < 16; ++i) {
        float x0, x1, p0, p1;
        unpack_f32x2(ffma2(pack_f32x2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), c2, link_b), x0, x1);
        if ((i & 3) == 3) exp2_poly2(x0, x1, p0, p1);
        else { p0 = ex2_approx(x0); p1 = ex2_approx(x1); }
        acc_b = fadd2(acc_b, pack_f32x2(p0, p1));
        pk[i] = pack_bf16(p0, p1);
    }
}

__global__ void __launch_bounds__(ATT_THREADS, 2)
attention_kernel(const __grid_constant__ CUtensorMap tmQKV, AttnParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sQ = smem;
    uint8_t* sK = smem + ATT_TILE_BYTES;
    uint8_t* sV = sK + ATT_STAGES * ATT_TILE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + ATT_STAGES * ATT_TILE_BYTES);
    uint64_t* q_full = bars;
    uint64_t* kv_full = bars + 1;
    uint64_t* kv_empty = kv_full + ATT_STAGES;
    uint64_t* s_full = kv_empty + ATT_STAGES;             // S(j) is in TMEM
    uint64_t* s_free = s_full + 1;                        // the softmax warps hold S(j) in registers: S(j+1) may be issued
    uint64_t* p_ready = s_free + 1;                       // P(j) is in TMEM
    uint64_t* pv_done = p_ready + 1;                      // O += P(j) V_j has retired
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int head = p.reverse ? gridDim.y - 1 - blockIdx.y : blockIdx.y;
    const int copy = p.reverse ? gridDim.z - 1 - blockIdx.z : blockIdx.z;
    const int q0 = blockIdx.x * ATT_TILE;
    const int nkv = (p.tokens + ATT_TILE - 1) / ATT_TILE;
    const int hidden = p.heads * ATT_HD;

    if (warp == ATT_W_TMA && elect_one()) {
        tma_prefetch_desc(&tmQKV);
        mbar_init(q_full, 1);
        for (int s = 0; s < ATT_STAGES; ++s) {
            mbar_init(&kv_full[s], 1);
            mbar_init(&kv_empty[s], 1);
        }
        mbar_init(s_full, 1);
        mbar_init(s_free, 4);
        mbar_init(p_ready, 4);
        mbar_init(pv_done, 1);
        fence_barrier_init();
    }
    if (warp == ATT_W_MMA) tmem_alloc<ATT_TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == ATT_W_TMA) {
        if (elect_one()) {          // elect.sync: ptxas emits straight-line UTMALDG / UTCHMMA (no per-lane BRA.U.ANY loop)
            mbar_expect_tx(q_full, ATT_TILE_BYTES);
            tma_load_3d(sQ, &tmQKV, q_full, head * ATT_HD, q0, copy);
            for (int j = 0; j < nkv; ++j) {
                const int st = j % ATT_STAGES;
                mbar_wait(&kv_empty[st], ((j / ATT_STAGES) & 1) ^ 1);
                mbar_expect_tx(&kv_full[st], 2 * ATT_TILE_BYTES);
                tma_load_3d(sK + st * ATT_TILE_BYTES, &tmQKV, &kv_full[st], hidden + head * ATT_HD, j * ATT_TILE, copy);
                tma_load_3d(sV + st * ATT_TILE_BYTES, &tmQKV, &kv_full[st], 2 * hidden + head * ATT_HD, j * ATT_TILE, copy);
            }
        }
    } else if (warp == ATT_W_MMA) {
        if (elect_one()) {
            constexpr uint32_t idesc_pv = make_idesc_bf16(ATT_TILE, ATT_HD, true);
            const uint32_t tS = tmem_base + ATT_S_COL, tO = tmem_base + ATT_O_COL, tP = tmem_base + ATT_P_COL;
            const uint64_t q_desc = make_smem_desc_sw128(smem_u32(sQ), 16, 1024);
            const uint64_t k_desc0 = make_smem_desc_sw128(smem_u32(sK), 16, 1024);
            const uint64_t v_desc0 = make_smem_desc_sw128(smem_u32(sV), 16384, 1024);
            auto issue_s = [&](int j) {                   // S = Q K_j^T
                const int nk = min(ATT_TILE, p.tokens - j * ATT_TILE);
                const uint32_t idesc_s = make_idesc_bf16(ATT_TILE, nk, false);
                const uint64_t kd = k_desc0 + static_cast<uint64_t>((j % ATT_STAGES) * (ATT_TILE_BYTES >> 4));
#pragma unroll
                for (int k = 0; k < ATT_HD / 16; ++k) umma_ss(tS, q_desc + 2 * k, kd + 2 * k, idesc_s, k != 0 ? 1u : 0u);
                umma_commit(s_full);
            };
            auto issue_pv = [&](int j) {                  // O += P V_j
                const int nk = min(ATT_TILE, p.tokens - j * ATT_TILE);
                const uint64_t vd = v_desc0 + static_cast<uint64_t>((j % ATT_STAGES) * (ATT_TILE_BYTES >> 4));
                if (nk == ATT_TILE) {
#pragma unroll
                    for (int ks = 0; ks < ATT_TILE / 16; ++ks) umma_ts(tO, tP + ks * 8, vd + 128 * ks, idesc_pv, (j | ks) != 0 ? 1u : 0u);
                } else {
                    for (int ks = 0; ks < nk / 16; ++ks) umma_ts(tO, tP + ks * 8, vd + 128 * ks, idesc_pv, (j | ks) != 0 ? 1u : 0u);
                }
                umma_commit(pv_done);
                umma_commit(&kv_empty[j % ATT_STAGES]);   // done with K_j / V_j
            };
            mbar_wait(q_full, 0);
            mbar_wait(&kv_full[0], 0);
            tc_fence_after();
            issue_s(0);
            for (int j = 0; j < nkv; ++j) {
                if (j + 1 < nkv) {
                    mbar_wait(&kv_full[(j + 1) % ATT_STAGES], ((j + 1) / ATT_STAGES) & 1);
                    mbar_wait(s_free, j & 1);             // S(j) sits in the softmax warps' registers
                    tc_fence_after();
                    issue_s(j + 1);
                }
                mbar_wait(p_ready, j & 1);
                tc_fence_after();
                issue_pv(j);
            }
        }
    } else {
        const int row = warp * 32 + lane;
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
        const uint32_t tS = t_lane + ATT_S_COL, tO = t_lane + ATT_O_COL, tP = t_lane + ATT_P_COL;
        const float c = p.scale_log2;
        const uint64_t c2 = pack_f32x2(c, c);
        const uint64_t zero2 = pack_f32x2(p.zero, p.zero);
        float m_ref = -INFINITY;
        uint64_t l2 = 0ull, l2b = 0ull;                   // running row sum, four partial sums
        uint32_t r[ATT_TILE];
        for (int j = 0; j < nkv; ++j) {
            const int nk = min(ATT_TILE, p.tokens - j * ATT_TILE);
            mbar_wait(s_full, j & 1);
            tc_fence_after();
            // the whole score row into registers, then hand the S buffer back to the tensor pipe at once
            if (nk == ATT_TILE) {
                tmem_ld32(tS, r); tmem_ld32(tS + 32, r + 32); tmem_ld32(tS + 64, r + 64); tmem_ld32(tS + 96, r + 96);
            } else {
#pragma unroll
                for (int col = 0; col < ATT_TILE; col += 16) {
                    if (col < nk) {
                        tmem_ld16(tS + col, r + col);
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i) r[col + i] = 0xff800000u;   // -inf: exp2 -> 0, never read by P.V
                    }
                }
            }
            tmem_wait_ld();
            tc_fence_before();
            __syncwarp();
            if (elect_one()) mbar_arrive(s_free);
            float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
            for (int i = 0; i < ATT_TILE; i += 8) {
                m0 = fmax3(m0, __uint_as_float(r[i]), __uint_as_float(r[i + 1]));
                m1 = fmax3(m1, __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
                m2 = fmax3(m2, __uint_as_float(r[i + 4]), __uint_as_float(r[i + 5]));
                m3 = fmax3(m3, __uint_as_float(r[i + 6]), __uint_as_float(r[i + 7]));
            }
            const float mt = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
            if (j == 0) m_ref = mt;
            // lazy rescaling: the reference max is replaced (and O, l rescaled) only when this tile exceeds it by 2^8
            const bool need = (mt - m_ref) * c > ATT_RESCALE_LOG2;
            bool pv_waited = false;
            if (__any_sync(0xffffffffu, need)) {
                if (j > 0) { mbar_wait(pv_done, (j - 1) & 1); tc_fence_after(); pv_waited = true; }   // O is quiescent
                const float m_new = fmaxf(m_ref, mt);
                const float sc = ex2_approx((m_ref - m_new) * c);
                l2 = ffma2(l2, pack_f32x2(sc, sc), 0ull);
                l2b = ffma2(l2b, pack_f32x2(sc, sc), 0ull);
#pragma unroll
                for (int cidx = 0; cidx < ATT_HD; cidx += 16) {
                    uint32_t o[16];
                    tmem_ld16(tO + cidx, o);
                    tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * sc);
                    tmem_st16(tO + cidx, o);
                }
                m_ref = m_new;
            }
            const float mc = m_ref * c;
            const uint64_t nmc2 = pack_f32x2(-mc, -mc);
            uint32_t pk[16];
            exp_chunk(r, pk, c2, nmc2, zero2, l2, l2b);
            // P may only be overwritten once P(j-1) . V has retired (checked here, a quarter of the pass later)
            if (j > 0 && !pv_waited) { mbar_wait(pv_done, (j - 1) & 1); tc_fence_after(); }
            tmem_st16(tP, pk);
            exp_chunk(r + 32, pk, c2, nmc2, zero2, l2, l2b);
            tmem_st16(tP + 16, pk);
            exp_chunk(r + 64, pk, c2, nmc2, zero2, l2, l2b);
            tmem_st16(tP + 32, pk);
            exp_chunk(r + 96, pk, c2, nmc2, zero2, l2, l2b);
            tmem_st16(tP + 48, pk);
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (elect_one()) mbar_arrive(p_ready);
        }
        mbar_wait(pv_done, (nkv - 1) & 1);
        tc_fence_after();
        const int q = q0 + row;
        float la, lb;
        unpack_f32x2(fadd2(l2, l2b), la, lb);
        const float inv = 1.0f / (la + lb);
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

    tc_fence_before();
    __syncthreads();
    if (warp == ATT_W_MMA) tmem_dealloc<ATT_TMEM_COLS>(tmem_base);
}

}  // namespace b200x

using namespace b200x;

extern "C" int b200x_attention(const void* d_qkv, void* d_out, int copies, int tokens, int heads, int head_dim, int reverse,
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
    AttnParams p{tokens, heads, reinterpret_cast<__nv_bfloat16*>(d_out), 0.125f * 1.4426950408889634f, 0.0f, reverse != 0};
    B200X_TRY(ensure_kernel_smem(reinterpret_cast<const void*>(attention_kernel), ATT_SMEM, true));
    dim3 grid(ceil_div(tokens, ATT_TILE), heads, copies);
    attention_kernel<<<grid, ATT_THREADS, ATT_SMEM, static_cast<cudaStream_t>(stream)>>>(tm, p);
    B200X_CUDA_TRY(cudaGetLastError());
    return B200X_OK;
}
