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

// NQ = query tiles per CTA.  NQ = 2: one CTA per SM, two tiles ping-pong and share K / V.  NQ = 1: two CTAs per SM (half the
// shared memory, TMEM columns and registers each); the co-resident CTAs run free of each other, so one CTA's load / max /
// barrier phases and its prologue / epilogue fall into the other's exponential phase instead of lining up with it.
// warps [0, 4 NQ): softmax (4 per tile), warp 4 NQ: TMA, warps 4 NQ + 1 ...: one MMA issuer per tile.
// The producer / issuer roles sit on the HIGHEST warp ids: the SM's warp arbiter prefers higher ids, and an issuer that
// loses its issue slots to the softmax warps delays every tcgen05.mma by hundreds of cycles.
constexpr int ATT_TILE = 128;
constexpr int ATT_HD = 64;
constexpr int ATT_TILE_BYTES = ATT_TILE * ATT_HD * 2;     // 16 KB
template <int NQ> struct AttCfg {
    static constexpr int THREADS = 32 * (5 * NQ + 1);
    static constexpr int W_TMA = 4 * NQ, W_MMA = 4 * NQ + 1;
    static constexpr int KV_STAGES = NQ == 2 ? 4 : 3;
    static constexpr int SMEM = ATT_TILE_BYTES * (NQ + 2 * KV_STAGES) + 256;     // dynamic shared memory is 1024-byte aligned
    static constexpr uint32_t TMEM_COLS = 256 * NQ;
    static constexpr uint32_t S_COL = 0, O_COL = 128 * NQ, P_COL = 192 * NQ;
};
constexpr float ATT_RESCALE_LOG2 = 8.0f;

struct AttnParams {
    int tokens;        // tokens per copy (multiple of 16)
    int heads;
    __nv_bfloat16* out;   // [copies * tokens, heads * 64]
    float scale_log2;     // (1/sqrt(64)) * log2(e)
    long long* prof;      // diagnostic variant 32: [cta][11 warps][4] cycle counters (else unused)
    float zero;           // always 0.0f: an operand ptxas cannot fold (see exp_chunk)
    int reverse;          // walk (copy, head) from the end: L2 reuse of the QKV GEMM's last output (see runtime.cu)
};

// 32 scores (registers r[0..32)) -> p = 2^(s*c - m_ref*c) -> bf16 pairs (round to nearest) -> 16 TMEM columns of P.
// Packed fp32x2 math keeps the issue cost at 2.5 slots per score: 1/2 FFMA2, MUFU, 1/2 F2FP, 1/2 FADD2.
// Scheduling: with the whole row in registers ptxas would hoist all 64 FFMA2 and then emit the MUFUs in one long run,
// which blocks the (in-order) warp on the MUFU queue while its other work waits.  The two 16-score halves of a chunk
// therefore take their addend from `link` = fma(sum two halves back, 0, -m*c): a true data dependence on older results
// (value unchanged) that keeps at most two halves in flight, so MUFU runs stay short and interleave with FMA work.
// which of the 8 score pairs of a half chunk take the FMA-pipe polynomial instead of the MUFU (bit 256: 2 of 8, 8192: 1 of 8,
// 16384: 3 of 8)
template <int DBG>
__device__ __forceinline__ constexpr bool att_poly_pair(int i) {
    return ((DBG & 256) && (i & 3) == 3) || ((DBG & 8192) && (i & 7) == 7) || ((DBG & 16384) && ((i & 7) == 2 || (i & 7) == 5 || (i & 7) == 7));
}
template <int DBG>
__device__ __forceinline__ void exp_chunk(const uint32_t* r, uint32_t (&pk)[16], uint64_t c2, uint64_t nmc2, uint64_t zero2,
                                          uint64_t& acc_a, uint64_t& acc_b) {
    const uint64_t link_a = ffma2(acc_a, zero2, nmc2);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float x0, x1;
        unpack_f32x2(ffma2(pack_f32x2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), c2, link_a), x0, x1);
        float p0, p1;
        if (att_poly_pair<DBG>(i)) exp2_poly2(x0, x1, p0, p1);
        else { p0 = (DBG & 1) ? x0 : ex2_approx(x0); p1 = (DBG & 1) ? x1 : ex2_approx(x1); }
        acc_a = fadd2(acc_a, pack_f32x2(p0, p1));
        pk[i] = pack_bf16(p0, p1);
    }
    const uint64_t link_b = ffma2(acc_b, zero2, nmc2);
#pragma unroll
    for (int i = 8; i < 16; ++i) {
        float x0, x1;
        unpack_f32x2(ffma2(pack_f32x2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), c2, link_b), x0, x1);
        float p0, p1;
        if (att_poly_pair<DBG>(i)) exp2_poly2(x0, x1, p0, p1);
        else { p0 = (DBG & 1) ? x0 : ex2_approx(x0); p1 = (DBG & 1) ? x1 : ex2_approx(x1); }
        acc_b = fadd2(acc_b, pack_f32x2(p0, p1));
        pk[i] = pack_bf16(p0, p1);
    }
}

// diagnostic variant 32: CTA (0,0,0) appends (event << 56 | step << 48 | clock) records per warp behind the counters
#define ATT_TRACE(ev, step) do { if ((DBG & 32) && trace != nullptr && tr_n < 120) { trace[tr_n++] = (static_cast<long long>(ev) << 56) | (static_cast<long long>(step) << 48) | (clock64() & 0xFFFFFFFFFFFFll); } } while (0)

template <int DBG, int NQ>
__global__ void __launch_bounds__(AttCfg<NQ>::THREADS, 3 - NQ)
attention_kernel(const __grid_constant__ CUtensorMap tmQKV, AttnParams p) {
    using Cfg = AttCfg<NQ>;
    constexpr int ATT_KV_STAGES = Cfg::KV_STAGES, ATT_W_TMA = Cfg::W_TMA, ATT_W_MMA = Cfg::W_MMA;
    constexpr uint32_t ATT_TMEM_COLS = Cfg::TMEM_COLS, ATT_S_COL = Cfg::S_COL, ATT_O_COL = Cfg::O_COL, ATT_P_COL = Cfg::P_COL;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sQ = smem;                                   // NQ tiles
    uint8_t* sK = smem + NQ * ATT_TILE_BYTES;
    uint8_t* sV = sK + ATT_KV_STAGES * ATT_TILE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + ATT_KV_STAGES * ATT_TILE_BYTES);
    uint64_t* q_full = bars;
    uint64_t* kv_full = bars + 1;
    uint64_t* kv_empty = kv_full + ATT_KV_STAGES;
    uint64_t* s_full = kv_empty + ATT_KV_STAGES;          // [2] S_x(j) is in TMEM
    uint64_t* s_free = s_full + 2;                        // [2] softmax x holds S_x(j) in registers: S_x(j+1) may be issued
    uint64_t* p_ready = s_free + 2;                       // [2] P_x(j) is in TMEM
    uint64_t* pv_done = p_ready + 2;                      // [2] O_x += P_x(j) V_j has retired
    uint64_t* turn = pv_done + 2;                         // [2] MUFU hand-over between the two softmax groups
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(turn + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int head = p.reverse ? gridDim.y - 1 - blockIdx.y : blockIdx.y, copy = p.reverse ? gridDim.z - 1 - blockIdx.z : blockIdx.z;
    const int q0 = blockIdx.x * NQ * ATT_TILE;
    const int nq = (NQ == 2 && q0 + ATT_TILE < p.tokens) ? 2 : 1;    // query tiles of this block that hold valid rows
    const int nkv = (p.tokens + ATT_TILE - 1) / ATT_TILE;
    const int hidden = p.heads * ATT_HD;
    long long* prof = nullptr;
    if ((DBG & 32)) prof = p.prof + ((static_cast<long long>(blockIdx.z) * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * (4 * Cfg::THREADS / 32) + warp * 4;
    long long* trace = nullptr;
    int tr_n = 0;
    if ((DBG & 32) && blockIdx.x == 1 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0)
        trace = p.prof + static_cast<long long>(gridDim.x) * gridDim.y * gridDim.z * (4 * Cfg::THREADS / 32) + warp * 128;

    if (warp == ATT_W_TMA && elect_one()) {
        tma_prefetch_desc(&tmQKV);
        mbar_init(q_full, 1);
        for (int s = 0; s < ATT_KV_STAGES; ++s) {
            mbar_init(&kv_full[s], 1);
            mbar_init(&kv_empty[s], nq);
        }
        for (int x = 0; x < NQ; ++x) {
            mbar_init(&s_full[x], 1);
            mbar_init(&s_free[x], 4);
            mbar_init(&p_ready[x], 4);
            mbar_init(&pv_done[x], 1);
            mbar_init(&turn[x], 4);
        }
        fence_barrier_init();
    }
    if (warp == ATT_W_MMA) tmem_alloc<ATT_TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == ATT_W_TMA) {
        if (!(DBG & 1024) && elect_one()) {          // elect.sync: ptxas emits straight-line UTMALDG / UTCHMMA (no per-lane BRA.U.ANY loop)
            if ((DBG & 32) && blockIdx.x == 1 && blockIdx.y == 0 && blockIdx.z == 0)
                trace = p.prof + static_cast<long long>(gridDim.x) * gridDim.y * gridDim.z * (4 * Cfg::THREADS / 32) + warp * 128;
            ATT_TRACE(0, 0);
            mbar_expect_tx(q_full, nq * ATT_TILE_BYTES);
            tma_load_3d(sQ, &tmQKV, q_full, head * ATT_HD, q0, copy);
            if (nq == 2) tma_load_3d(sQ + ATT_TILE_BYTES, &tmQKV, q_full, head * ATT_HD, q0 + ATT_TILE, copy);
            for (int j = 0; j < nkv; ++j) {
                const int st = j % ATT_KV_STAGES;
                const uint32_t ph = (j / ATT_KV_STAGES) & 1;
                mbar_wait(&kv_empty[st], ph ^ 1);
                ATT_TRACE(1, j);
                mbar_expect_tx(&kv_full[st], 2 * ATT_TILE_BYTES);
                tma_load_3d(sK + st * ATT_TILE_BYTES, &tmQKV, &kv_full[st], hidden + head * ATT_HD, j * ATT_TILE, copy);
                tma_load_3d(sV + st * ATT_TILE_BYTES, &tmQKV, &kv_full[st], 2 * hidden + head * ATT_HD, j * ATT_TILE, copy);
            }
        }
    } else if (warp >= ATT_W_MMA) {
        // one issuer per query tile: the tiles' chains (S -> registers -> next S;  P -> P.V) stay independent of each other
        const int x = warp - ATT_W_MMA;
        if (!(DBG & 1024) && x < nq && elect_one()) {
            if ((DBG & 32) && blockIdx.x == 1 && blockIdx.y == 0 && blockIdx.z == 0)
                trace = p.prof + static_cast<long long>(gridDim.x) * gridDim.y * gridDim.z * (4 * Cfg::THREADS / 32) + warp * 128;
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
                if (!(DBG & 16)) {
#pragma unroll
                    for (int k = 0; k < ATT_HD / 16; ++k) umma_ss(tS, q_desc + 2 * k, kd + 2 * k, idesc_s, k != 0 ? 1u : 0u);
                }
                umma_commit(&s_full[x]);
            };
            auto issue_pv = [&](int j) {                  // O_x += P_x V_j
                const int nk = min(ATT_TILE, p.tokens - j * ATT_TILE);
                const uint64_t vd = v_desc0 + static_cast<uint64_t>((j % ATT_KV_STAGES) * (ATT_TILE_BYTES >> 4));
                if (!(DBG & 8)) {
                    if (nk == ATT_TILE) {
#pragma unroll
                        for (int ks = 0; ks < ATT_TILE / 16; ++ks) umma_ts(tO, tP + ks * 8, vd + 128 * ks, idesc_pv, (j | ks) != 0 ? 1u : 0u);
                    } else {
                        for (int ks = 0; ks < nk / 16; ++ks) umma_ts(tO, tP + ks * 8, vd + 128 * ks, idesc_pv, (j | ks) != 0 ? 1u : 0u);
                    }
                }
                umma_commit(&pv_done[x]);
                umma_commit(&kv_empty[j % ATT_KV_STAGES]);              // this tile is done with K_j / V_j
            };
            long long pc_kv = 0, pc_p = 0, pc_f = 0, pc_t = 0, pc_start = 0;
            if ((DBG & 32)) pc_start = clock64();
            mbar_wait(q_full, 0);
            mbar_wait(&kv_full[0], 0);
            tc_fence_after();
            issue_s(0);
            for (int j = 0; j < nkv; ++j) {
                if (j + 1 < nkv) {
                    if ((DBG & 32)) pc_t = clock64();
                    mbar_wait(&kv_full[(j + 1) % ATT_KV_STAGES], ((j + 1) / ATT_KV_STAGES) & 1);
                    if ((DBG & 32)) { const long long t = clock64(); pc_kv += t - pc_t; pc_t = t; }
                    ATT_TRACE(1, j);
                    if (DBG & 32768) mbar_spin_wait(&s_free[x], j & 1); else
                    mbar_wait(&s_free[x], j & 1);                       // S_x(j) sits in the softmax warps' registers
                    tc_fence_after();
                    if ((DBG & 32)) pc_f += clock64() - pc_t;
                    ATT_TRACE(2, j);
                    issue_s(j + 1);
                    ATT_TRACE(4, j);
                }
                if ((DBG & 32)) pc_t = clock64();
                if (DBG & 32768) mbar_spin_wait(&p_ready[x], j & 1); else
                mbar_wait(&p_ready[x], j & 1);
                tc_fence_after();
                if ((DBG & 32)) pc_p += clock64() - pc_t;
                ATT_TRACE(3, j);
                issue_pv(j);
                ATT_TRACE(5, j);
            }
            if ((DBG & 32)) { prof[0] = pc_kv; prof[1] = pc_p; prof[2] = clock64() - pc_start; prof[3] = pc_f; }
        }
    } else {
        const int x = warp >> 2;                          // query tile of this softmax warp group (warps 0-3: A, 4-7: B)
        if (x < nq) {
            const int quarter = warp & 3;
            const int row = quarter * 32 + lane;
            const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
            const uint32_t tS = t_lane + ATT_S_COL + x * ATT_TILE;
            const uint32_t tO = t_lane + ATT_O_COL + x * ATT_HD;
            const uint32_t tP = t_lane + ATT_P_COL + x * ATT_HD;
            const float c = p.scale_log2;
            const uint64_t c2 = pack_f32x2(c, c);
            const uint64_t zero2 = pack_f32x2(p.zero, p.zero);
            float m_ref = -INFINITY;
            uint64_t l2 = 0ull, l2b = 0ull;               // running row sum, four partial sums
            long long pc_wait = 0, pc_pass = 0, pc_tail = 0, pc_t = 0, pc_ld = 0, pc_max = 0, pc_pv = 0, pc_u = 0;
            constexpr bool PIPE = (DBG & 65536) != 0;         // S(j+1) is loaded while the P(j) stores drain (see attention_fwd_kernel)
            uint32_t r[ATT_TILE];
            for (int j = 0; j < nkv; ++j) {
                const int nk = min(ATT_TILE, p.tokens - j * ATT_TILE);
                if ((DBG & 32)) pc_t = clock64();
                if (!PIPE || j == 0) {
                if (!(DBG & 1024)) mbar_wait(&s_full[x], j & 1);
                tc_fence_after();
                }
                if ((DBG & 32)) { const long long t = clock64(); pc_wait += t - pc_t; pc_t = t; }
                ATT_TRACE(1, j);
                // the whole score row into registers, then hand the S buffer back to the tensor pipe at once
                if (PIPE && j > 0) {
                    // already loaded and released at the end of the previous iteration
                } else if ((DBG & 3072) == 1024) {
#pragma unroll
                    for (int i = 0; i < ATT_TILE; ++i) r[i] = __float_as_uint(p.zero * (i + j) - 0.01f * (i + (lane & 7)));
                } else if ((DBG & 7) != 4) {
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
                }
                if ((DBG & 32)) { pc_u = clock64(); pc_ld += pc_u - pc_t; }
                ATT_TRACE(2, j);
                if (!PIPE || j == 0) {
                tc_fence_before();
                __syncwarp();
                if (!(DBG & 1024) && elect_one()) mbar_arrive(&s_free[x]);
                }
                if ((DBG & 7) != 4) {
                    float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
                    if (!(DBG & 4096) || j == 0)
#pragma unroll
                    for (int i = 0; i < ATT_TILE; i += 8) {
                        m0 = fmax3(m0, __uint_as_float(r[i]), __uint_as_float(r[i + 1]));
                        m1 = fmax3(m1, __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
                        m2 = fmax3(m2, __uint_as_float(r[i + 4]), __uint_as_float(r[i + 5]));
                        m3 = fmax3(m3, __uint_as_float(r[i + 6]), __uint_as_float(r[i + 7]));
                    }
                    float mt = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
                    if (j == 0) m_ref = mt;
                    if ((DBG & 4096) && j > 0) mt = m_ref;
                    if ((DBG & 32)) { const long long t = clock64(); pc_max += t - pc_u; pc_u = t; }
                    ATT_TRACE(3, j);
                    // lazy rescaling: the reference max is replaced (and O, l rescaled) only when this tile exceeds it by 2^8
                    const bool need = (mt - m_ref) * c > ATT_RESCALE_LOG2;
                    bool pv_waited = false;
                    if (__any_sync(0xffffffffu, need)) {
                        if (j > 0 && !(DBG & 1024)) { mbar_wait(&pv_done[x], (j - 1) & 1); tc_fence_after(); pv_waited = true; }   // O_x is quiescent
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
                    // one-time stagger: tile B starts its first exponential phase when tile A has finished its own, so that
                    // from then on one group's load / max / wait phases fall into the other group's MUFU phase
                    // strict alternation of the exponential phases (both tiles valid): group x runs its MUFU phase j after
                    // the other group has finished its phase j (x = 1) / j - 1 (x = 0)
                    if ((DBG & 64) && nq == 2) {
                        if (x == 1) mbar_wait(&turn[0], j & 1);
                        else if (j > 0) mbar_wait(&turn[1], (j - 1) & 1);
                    }
                    if ((DBG & 128) && nq == 2 && x == 1 && j == 0) mbar_wait(&turn[0], 0);
                    uint32_t pk[16];
                    exp_chunk<DBG>(r, pk, c2, nmc2, zero2, l2, l2b);
                    if ((DBG & 32)) pc_u = clock64();
                    // P_x may only be overwritten once P_x(j-1) . V has retired (checked here, a quarter of the pass later)
                    if (j > 0 && !pv_waited && !(DBG & 1024)) { mbar_wait(&pv_done[x], (j - 1) & 1); tc_fence_after(); }
                    if ((DBG & 32)) pc_pv += clock64() - pc_u;
                    ATT_TRACE(4, j);
                    if ((DBG & 32) && j + 1 < nkv) { const bool rdy = mbar_try_wait(&s_full[x], (j + 1) & 1); ATT_TRACE(rdy ? 8 : 7, j); }
                    if (!(DBG & 512) && (DBG & 3072) != 1024) tmem_st16(tP, pk); else asm volatile("" :: "r"(pk[0] ^ pk[5] ^ pk[9] ^ pk[15]));
                    exp_chunk<DBG>(r + 32, pk, c2, nmc2, zero2, l2, l2b);
                    if (!(DBG & 512) && (DBG & 3072) != 1024) tmem_st16(tP + 16, pk); else asm volatile("" :: "r"(pk[0] ^ pk[5] ^ pk[9] ^ pk[15]));
                    exp_chunk<DBG>(r + 64, pk, c2, nmc2, zero2, l2, l2b);
                    if (!(DBG & 512) && (DBG & 3072) != 1024) tmem_st16(tP + 32, pk); else asm volatile("" :: "r"(pk[0] ^ pk[5] ^ pk[9] ^ pk[15]));
                    exp_chunk<DBG>(r + 96, pk, c2, nmc2, zero2, l2, l2b);
                    if (!(DBG & 512) && (DBG & 3072) != 1024) tmem_st16(tP + 48, pk); else asm volatile("" :: "r"(pk[0] ^ pk[5] ^ pk[9] ^ pk[15]));
                    if (((DBG & 64) && nq == 2) || ((DBG & 128) && nq == 2 && x == 0 && j == 0)) {
                        __syncwarp();
                        if (elect_one()) mbar_arrive(&turn[x]);
                    }
                } else {
                    if (j > 0) { mbar_wait(&pv_done[x], (j - 1) & 1); tc_fence_after(); }
                    l2 = pack_f32x2(1.f, 0.f);
                }
                if ((DBG & 32)) { const long long t = clock64(); pc_pass += t - pc_t; pc_t = t; }
                ATT_TRACE(5, j);
                if ((DBG & 32) && j + 1 < nkv) { const bool rdy = mbar_try_wait(&s_full[x], (j + 1) & 1); ATT_TRACE(rdy ? 10 : 9, j); }
                if (PIPE && j + 1 < nkv) {
                    mbar_wait(&s_full[x], (j + 1) & 1);
                    tc_fence_after();
                    const int nk1 = min(ATT_TILE, p.tokens - (j + 1) * ATT_TILE);
                    if (nk1 == ATT_TILE) {
                        tmem_ld32(tS, r); tmem_ld32(tS + 32, r + 32); tmem_ld32(tS + 64, r + 64); tmem_ld32(tS + 96, r + 96);
                    } else {
#pragma unroll
                        for (int col = 0; col < ATT_TILE; col += 16) {
                            if (col < nk1) {
                                tmem_ld16(tS + col, r + col);
                            } else {
#pragma unroll
                                for (int i = 0; i < 16; ++i) r[col + i] = 0xff800000u;
                            }
                        }
                    }
                }
                tmem_wait_st();
                tc_fence_before();
                __syncwarp();
                if (!(DBG & 1024) && elect_one()) mbar_arrive(&p_ready[x]);
                if (PIPE && j + 1 < nkv) {
                    tmem_wait_ld();
                    tc_fence_before();
                    __syncwarp();
                    if (elect_one()) mbar_arrive(&s_free[x]);
                }
                if ((DBG & 32)) { const long long t = clock64(); pc_tail += t - pc_t; pc_t = t; }
                ATT_TRACE(6, j);
            }
            if ((DBG & 32) && lane == 0) { prof[0] = pc_wait; prof[1] = pc_pass; prof[2] = pc_tail; prof[3] = (pc_ld << 42) | (pc_max << 21) | pc_pv; }
            if (!(DBG & 1024)) mbar_wait(&pv_done[x], (nkv - 1) & 1);
            tc_fence_after();
            const int q = q0 + x * ATT_TILE + row;
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
    }

    tc_fence_before();
    __syncthreads();
    if (warp == ATT_W_MMA) tmem_dealloc<ATT_TMEM_COLS>(tmem_base);
}



// ----------------------------------------------------------------------------- production kernel
// attention_kernel<256, 1> (one 128-row query tile per CTA, two CTAs per SM, one row per softmax thread, a quarter of the
// exponentials on the FMA pipe) without the diagnostic plumbing and with the softmax loop software-pipelined across key
// tiles: the tcgen05.ld of S(j+1) is issued BEFORE the wait on the P(j) stores, so the TMEM read latency (~400 cycles per
// tile in the traces) hides behind the store drain (~500 cycles) instead of following it.
//   warps 0-3: softmax;  warp 4: TMA producer;  warp 5: TMEM allocator + tcgen05.mma issuer
// TMEM columns: S [0,128)  O [128,192)  P [192,256).
template <bool POLY>
__global__ void __launch_bounds__(AttCfg<1>::THREADS, 2)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, AttnParams p) {
    using Cfg = AttCfg<1>;
    constexpr int STAGES = Cfg::KV_STAGES;
    constexpr int VARIANT = POLY ? 256 : 0;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sQ = smem;
    uint8_t* sK = smem + ATT_TILE_BYTES;
    uint8_t* sV = sK + STAGES * ATT_TILE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + STAGES * ATT_TILE_BYTES);
    uint64_t* q_full = bars;
    uint64_t* kv_full = bars + 1;
    uint64_t* kv_empty = kv_full + STAGES;
    uint64_t* s_full = kv_empty + STAGES;
    uint64_t* s_free = s_full + 1;
    uint64_t* p_ready = s_free + 1;
    uint64_t* pv_done = p_ready + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int head = blockIdx.y, copy = blockIdx.z;
    const int q0 = blockIdx.x * ATT_TILE;
    const int nkv = (p.tokens + ATT_TILE - 1) / ATT_TILE;
    const int hidden = p.heads * ATT_HD;

    if (warp == Cfg::W_TMA && elect_one()) {
        tma_prefetch_desc(&tmQKV);
        mbar_init(q_full, 1);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&kv_full[s], 1);
            mbar_init(&kv_empty[s], 1);
        }
        mbar_init(s_full, 1);
        mbar_init(s_free, 4);
        mbar_init(p_ready, 4);
        mbar_init(pv_done, 1);
        fence_barrier_init();
    }
    if (warp == Cfg::W_MMA) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == Cfg::W_TMA) {
        if (elect_one()) {
            mbar_expect_tx(q_full, ATT_TILE_BYTES);
            tma_load_3d(sQ, &tmQKV, q_full, head * ATT_HD, q0, copy);
            for (int j = 0; j < nkv; ++j) {
                const int st = j % STAGES;
                mbar_wait(&kv_empty[st], ((j / STAGES) & 1) ^ 1);
                mbar_expect_tx(&kv_full[st], 2 * ATT_TILE_BYTES);
                tma_load_3d(sK + st * ATT_TILE_BYTES, &tmQKV, &kv_full[st], hidden + head * ATT_HD, j * ATT_TILE, copy);
                tma_load_3d(sV + st * ATT_TILE_BYTES, &tmQKV, &kv_full[st], 2 * hidden + head * ATT_HD, j * ATT_TILE, copy);
            }
        }
    } else if (warp == Cfg::W_MMA) {
        if (elect_one()) {
            constexpr uint32_t idesc_pv = make_idesc_bf16(ATT_TILE, ATT_HD, true);
            const uint32_t tS = tmem_base + Cfg::S_COL, tO = tmem_base + Cfg::O_COL, tP = tmem_base + Cfg::P_COL;
            const uint64_t q_desc = make_smem_desc_sw128(smem_u32(sQ), 16, 1024);
            const uint64_t k_desc0 = make_smem_desc_sw128(smem_u32(sK), 16, 1024);
            const uint64_t v_desc0 = make_smem_desc_sw128(smem_u32(sV), 16384, 1024);
            auto issue_s = [&](int j) {                   // S = Q K_j^T
                const int nk = min(ATT_TILE, p.tokens - j * ATT_TILE);
                const uint32_t idesc_s = make_idesc_bf16(ATT_TILE, nk, false);
                const uint64_t kd = k_desc0 + static_cast<uint64_t>((j % STAGES) * (ATT_TILE_BYTES >> 4));
#pragma unroll
                for (int k = 0; k < ATT_HD / 16; ++k) umma_ss(tS, q_desc + 2 * k, kd + 2 * k, idesc_s, k != 0 ? 1u : 0u);
                umma_commit(s_full);
            };
            auto issue_pv = [&](int j) {                  // O += P V_j
                const int nk = min(ATT_TILE, p.tokens - j * ATT_TILE);
                const uint64_t vd = v_desc0 + static_cast<uint64_t>((j % STAGES) * (ATT_TILE_BYTES >> 4));
                if (nk == ATT_TILE) {
#pragma unroll
                    for (int ks = 0; ks < ATT_TILE / 16; ++ks) umma_ts(tO, tP + ks * 8, vd + 128 * ks, idesc_pv, (j | ks) != 0 ? 1u : 0u);
                } else {
                    for (int ks = 0; ks < nk / 16; ++ks) umma_ts(tO, tP + ks * 8, vd + 128 * ks, idesc_pv, (j | ks) != 0 ? 1u : 0u);
                }
                umma_commit(pv_done);
                umma_commit(&kv_empty[j % STAGES]);
            };
            mbar_wait(q_full, 0);
            mbar_wait(&kv_full[0], 0);
            tc_fence_after();
            issue_s(0);
            for (int j = 0; j < nkv; ++j) {
                if (j + 1 < nkv) {
                    mbar_wait(&kv_full[(j + 1) % STAGES], ((j + 1) / STAGES) & 1);
                    mbar_wait(s_free, j & 1);
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
        const uint32_t tS = t_lane + Cfg::S_COL, tO = t_lane + Cfg::O_COL, tP = t_lane + Cfg::P_COL;
        const float c = p.scale_log2;
        const uint64_t c2 = pack_f32x2(c, c);
        const uint64_t zero2 = pack_f32x2(p.zero, p.zero);
        float m_ref = -INFINITY;
        uint64_t l2 = 0ull, l2b = 0ull;
        uint32_t r[ATT_TILE];
        // issue (not wait for) the TMEM loads of score tile j; columns past the last key read as -inf
        auto load_scores = [&](int j) {
            const int nk = min(ATT_TILE, p.tokens - j * ATT_TILE);
            if (nk == ATT_TILE) {
                tmem_ld32(tS, r); tmem_ld32(tS + 32, r + 32); tmem_ld32(tS + 64, r + 64); tmem_ld32(tS + 96, r + 96);
            } else {
#pragma unroll
                for (int col = 0; col < ATT_TILE; col += 16) {
                    if (col < nk) {
                        tmem_ld16(tS + col, r + col);
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i) r[col + i] = 0xff800000u;
                    }
                }
            }
        };
        mbar_wait(s_full, 0);
        tc_fence_after();
        load_scores(0);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (elect_one()) mbar_arrive(s_free);
        for (int j = 0; j < nkv; ++j) {
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
            exp_chunk<VARIANT>(r, pk, c2, nmc2, zero2, l2, l2b);
            if (j > 0 && !pv_waited) { mbar_wait(pv_done, (j - 1) & 1); tc_fence_after(); }   // P(j-1) . V has retired
            tmem_st16(tP, pk);
            exp_chunk<VARIANT>(r + 32, pk, c2, nmc2, zero2, l2, l2b);
            tmem_st16(tP + 16, pk);
            exp_chunk<VARIANT>(r + 64, pk, c2, nmc2, zero2, l2, l2b);
            tmem_st16(tP + 32, pk);
            exp_chunk<VARIANT>(r + 96, pk, c2, nmc2, zero2, l2, l2b);
            tmem_st16(tP + 48, pk);
            const bool more = j + 1 < nkv;
            if (more) {                                   // S(j+1) was issued when S(j) was released: it is (almost always) there
                mbar_wait(s_full, (j + 1) & 1);
                tc_fence_after();
                load_scores(j + 1);                       // in flight while the P stores drain
            }
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (elect_one()) mbar_arrive(p_ready);
            if (more) {
                tmem_wait_ld();
                tc_fence_before();
                __syncwarp();
                if (elect_one()) mbar_arrive(s_free);
            }
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
    if (warp == Cfg::W_MMA) tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
}

// ----------------------------------------------------------------------------- 64-key / in-place-P kernel (measured variant)
// THREE CTAs per SM instead of two: key tiles of 64, the bf16 P tile written over the score tile it was computed from
// (S / P 64 columns + O 64 columns = 128 TMEM columns per CTA), 64 KB of shared memory and <= 112 registers per thread.
// The chain of one CTA is strictly serial - S(j) -> softmax(j) -> P(j).V -> S(j+1), all ordered by the in-order tensor pipe,
// so neither an "S free" nor a "P.V done" barrier is needed - and the overlap comes from the other two resident CTAs.
//   warps 0-3: softmax (one query row per thread);  warp 4: TMA producer;  warp 5: TMEM allocator + tcgen05.mma issuer
struct AttK64 {
    static constexpr int THREADS = 192, W_TMA = 4, W_MMA = 5, KV = 64, KV_STAGES = 3;
    static constexpr int KV_BYTES = KV * ATT_HD * 2;                       // 8 KB per K or V tile
    static constexpr int SMEM = ATT_TILE_BYTES + 2 * KV_STAGES * KV_BYTES + 256;
    static constexpr uint32_t TMEM_COLS = 128, S_COL = 0, O_COL = 64;       // P aliases S: columns [0, 32)
};

template <bool POLY>
__global__ void __launch_bounds__(AttK64::THREADS, 3)
attention_k64_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, AttnParams p) {
    using Cfg = AttK64;
    constexpr int STAGES = Cfg::KV_STAGES, KV = Cfg::KV;
    constexpr int VARIANT = POLY ? 256 : 0;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sQ = smem;
    uint8_t* sK = smem + ATT_TILE_BYTES;
    uint8_t* sV = sK + STAGES * Cfg::KV_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + STAGES * Cfg::KV_BYTES);
    uint64_t* q_full = bars;
    uint64_t* kv_full = bars + 1;
    uint64_t* kv_empty = kv_full + STAGES;
    uint64_t* s_full = kv_empty + STAGES;
    uint64_t* p_ready = s_full + 1;
    uint64_t* pv_done = p_ready + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int head = p.reverse ? gridDim.y - 1 - blockIdx.y : blockIdx.y, copy = p.reverse ? gridDim.z - 1 - blockIdx.z : blockIdx.z;
    const int q0 = blockIdx.x * ATT_TILE;
    const int nkv = (p.tokens + KV - 1) / KV;
    const int hidden = p.heads * ATT_HD;

    if (warp == Cfg::W_TMA && elect_one()) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmKV);
        mbar_init(q_full, 1);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&kv_full[s], 1);
            mbar_init(&kv_empty[s], 1);
        }
        mbar_init(s_full, 1);
        mbar_init(p_ready, 4);
        mbar_init(pv_done, 1);
        fence_barrier_init();
    }
    if (warp == Cfg::W_MMA) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == Cfg::W_TMA) {
        if (elect_one()) {
            mbar_expect_tx(q_full, ATT_TILE_BYTES);
            tma_load_3d(sQ, &tmQ, q_full, head * ATT_HD, q0, copy);
            for (int j = 0; j < nkv; ++j) {
                const int st = j % STAGES;
                mbar_wait(&kv_empty[st], ((j / STAGES) & 1) ^ 1);
                mbar_expect_tx(&kv_full[st], 2 * Cfg::KV_BYTES);
                tma_load_3d(sK + st * Cfg::KV_BYTES, &tmKV, &kv_full[st], hidden + head * ATT_HD, j * KV, copy);
                tma_load_3d(sV + st * Cfg::KV_BYTES, &tmKV, &kv_full[st], 2 * hidden + head * ATT_HD, j * KV, copy);
            }
        }
    } else if (warp == Cfg::W_MMA) {
        if (elect_one()) {
            constexpr uint32_t idesc_pv = make_idesc_bf16(ATT_TILE, ATT_HD, true);
            const uint32_t tS = tmem_base + Cfg::S_COL, tO = tmem_base + Cfg::O_COL, tP = tmem_base + Cfg::S_COL;
            const uint64_t q_desc = make_smem_desc_sw128(smem_u32(sQ), 16, 1024);
            const uint64_t k_desc0 = make_smem_desc_sw128(smem_u32(sK), 16, 1024);
            const uint64_t v_desc0 = make_smem_desc_sw128(smem_u32(sV), 8192, 1024);
            mbar_wait(q_full, 0);
            for (int j = 0; j < nkv; ++j) {
                const int nk = min(KV, p.tokens - j * KV);
                const int st = j % STAGES;
                mbar_wait(&kv_full[st], (j / STAGES) & 1);
                tc_fence_after();
                // S(j) = Q K_j^T overwrites the columns P(j-1) was read from by the MMAs just before it (in-order pipe)
                const uint32_t idesc_s = make_idesc_bf16(ATT_TILE, nk, false);
                const uint64_t kd = k_desc0 + static_cast<uint64_t>(st * (Cfg::KV_BYTES >> 4));
#pragma unroll
                for (int k = 0; k < ATT_HD / 16; ++k) umma_ss(tS, q_desc + 2 * k, kd + 2 * k, idesc_s, k != 0 ? 1u : 0u);
                umma_commit(s_full);
                mbar_wait(p_ready, j & 1);
                tc_fence_after();
                const uint64_t vd = v_desc0 + static_cast<uint64_t>(st * (Cfg::KV_BYTES >> 4));
                for (int ks = 0; ks < nk / 16; ++ks) umma_ts(tO, tP + ks * 8, vd + 128 * ks, idesc_pv, (j | ks) != 0 ? 1u : 0u);
                umma_commit(pv_done);
                umma_commit(&kv_empty[st]);
            }
        }
    } else {
        const int row = warp * 32 + lane;
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
        const uint32_t tS = t_lane + Cfg::S_COL, tO = t_lane + Cfg::O_COL, tP = t_lane + Cfg::S_COL;
        const float c = p.scale_log2;
        const uint64_t c2 = pack_f32x2(c, c);
        const uint64_t zero2 = pack_f32x2(p.zero, p.zero);
        float m_ref = -INFINITY;
        uint64_t l2 = 0ull, l2b = 0ull;
        for (int j = 0; j < nkv; ++j) {
            const int nk = min(KV, p.tokens - j * KV);
            mbar_wait(s_full, j & 1);
            tc_fence_after();
            uint32_t r[KV];
            if (nk == KV) {
                tmem_ld32(tS, r);
                tmem_ld32(tS + 32, r + 32);
            } else {
#pragma unroll
                for (int col = 0; col < KV; col += 16) {
                    if (col < nk) {
                        tmem_ld16(tS + col, r + col);
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i) r[col + i] = 0xff800000u;
                    }
                }
            }
            tmem_wait_ld();
            float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
            for (int i = 0; i < KV; i += 8) {
                m0 = fmax3(m0, __uint_as_float(r[i]), __uint_as_float(r[i + 1]));
                m1 = fmax3(m1, __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
                m2 = fmax3(m2, __uint_as_float(r[i + 4]), __uint_as_float(r[i + 5]));
                m3 = fmax3(m3, __uint_as_float(r[i + 6]), __uint_as_float(r[i + 7]));
            }
            const float mt = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
            if (j == 0) m_ref = mt;
            const bool need = (mt - m_ref) * c > ATT_RESCALE_LOG2;
            if (__any_sync(0xffffffffu, need)) {          // O is quiescent: P(j-1).V retired before S(j) was written
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
            exp_chunk<VARIANT>(r, pk, c2, nmc2, zero2, l2, l2b);
            tmem_st16(tP, pk);                            // the whole score row sits in registers: its columns can be reused
            exp_chunk<VARIANT>(r + 32, pk, c2, nmc2, zero2, l2, l2b);
            tmem_st16(tP + 16, pk);
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
    if (warp == Cfg::W_MMA) tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
}

// ----------------------------------------------------------------------------- split-row kernel (measured variant)
// One 128-row query tile per CTA, two CTAs per SM, and TWO threads per query row: softmax warp w (0-7) owns TMEM lanes
// [32 (w & 3), +32) and the score columns [64 (w >> 2), +64) of every 128-key tile, so each SM sub-partition holds four
// softmax warps (two per CTA) with short phases instead of two with long ones - the MUFU sees exponential work from
// several warps at once and one warp's load / max / barrier latencies hide behind the others'.
//   warps 0-7 : softmax (row = 32 (w & 3) + lane, half = w >> 2);  warp 8: TMA producer;  warp 9: tcgen05.mma issuer
// The two threads of a row agree on the tile's row max through shared memory (bf16, 512 B) and one named barrier per
// row quarter; the running reference max is lazily replaced exactly as in the kernel above.  Each thread keeps the partial
// row sum of its own columns and normalises / stores its own 32 output columns; the partial sums meet once, at the end.
// TMEM columns: S [0,128)  O [128,192)  P [192,256).
struct AttSplit {
    static constexpr int THREADS = 320, W_TMA = 8, W_MMA = 9, KV_STAGES = 3;
    static constexpr int BAR_BYTES = 256, XCH_BYTES = 512;
    static constexpr int SMEM = ATT_TILE_BYTES * (1 + 2 * KV_STAGES) + BAR_BYTES + XCH_BYTES;
    static constexpr uint32_t TMEM_COLS = 256, S_COL = 0, O_COL = 128, P_COL = 192;
};

// rendezvous of the two warps that share a row quarter (named barriers 1-4, compile-time ids)
__device__ __forceinline__ void pair_bar_sync(int quarter) {
    switch (quarter) {
        case 0: asm volatile("bar.sync 1, 64;" ::: "memory"); break;
        case 1: asm volatile("bar.sync 2, 64;" ::: "memory"); break;
        case 2: asm volatile("bar.sync 3, 64;" ::: "memory"); break;
        default: asm volatile("bar.sync 4, 64;" ::: "memory"); break;
    }
}

template <int DBG>
__global__ void __launch_bounds__(AttSplit::THREADS, 2)
attention_split_kernel(const __grid_constant__ CUtensorMap tmQKV, AttnParams p) {
    using Cfg = AttSplit;
    constexpr int STAGES = Cfg::KV_STAGES;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sQ = smem;
    uint8_t* sK = smem + ATT_TILE_BYTES;
    uint8_t* sV = sK + STAGES * ATT_TILE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + STAGES * ATT_TILE_BYTES);
    uint64_t* q_full = bars;
    uint64_t* kv_full = bars + 1;
    uint64_t* kv_empty = kv_full + STAGES;
    uint64_t* s_full = kv_empty + STAGES;                 // S(j) is in TMEM
    uint64_t* s_free = s_full + 1;                        // the softmax warps hold S(j) in registers: S(j+1) may be issued
    uint64_t* p_ready = s_free + 1;                       // P(j) is in TMEM
    uint64_t* pv_done = p_ready + 1;                      // O += P(j) V_j has retired
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 1);
    __nv_bfloat16* xch = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(bars) + Cfg::BAR_BYTES);   // [2][128] tile row max per half

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int head = blockIdx.y, copy = blockIdx.z;
    const int q0 = blockIdx.x * ATT_TILE;
    const int nkv = (p.tokens + ATT_TILE - 1) / ATT_TILE;
    const int hidden = p.heads * ATT_HD;

    if (warp == Cfg::W_TMA && elect_one()) {
        tma_prefetch_desc(&tmQKV);
        mbar_init(q_full, 1);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&kv_full[s], 1);
            mbar_init(&kv_empty[s], 1);
        }
        mbar_init(s_full, 1);
        mbar_init(s_free, 8);
        mbar_init(p_ready, 8);
        mbar_init(pv_done, 1);
        fence_barrier_init();
    }
    if (warp == Cfg::W_MMA) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == Cfg::W_TMA) {
        if (elect_one()) {
            mbar_expect_tx(q_full, ATT_TILE_BYTES);
            tma_load_3d(sQ, &tmQKV, q_full, head * ATT_HD, q0, copy);
            for (int j = 0; j < nkv; ++j) {
                const int st = j % STAGES;
                mbar_wait(&kv_empty[st], ((j / STAGES) & 1) ^ 1);
                mbar_expect_tx(&kv_full[st], 2 * ATT_TILE_BYTES);
                tma_load_3d(sK + st * ATT_TILE_BYTES, &tmQKV, &kv_full[st], hidden + head * ATT_HD, j * ATT_TILE, copy);
                tma_load_3d(sV + st * ATT_TILE_BYTES, &tmQKV, &kv_full[st], 2 * hidden + head * ATT_HD, j * ATT_TILE, copy);
            }
        }
    } else if (warp == Cfg::W_MMA) {
        if (elect_one()) {
            constexpr uint32_t idesc_pv = make_idesc_bf16(ATT_TILE, ATT_HD, true);
            const uint32_t tS = tmem_base + Cfg::S_COL, tO = tmem_base + Cfg::O_COL, tP = tmem_base + Cfg::P_COL;
            const uint64_t q_desc = make_smem_desc_sw128(smem_u32(sQ), 16, 1024);
            const uint64_t k_desc0 = make_smem_desc_sw128(smem_u32(sK), 16, 1024);
            const uint64_t v_desc0 = make_smem_desc_sw128(smem_u32(sV), 16384, 1024);
            auto issue_s = [&](int j) {                   // S = Q K_j^T
                const int nk = min(ATT_TILE, p.tokens - j * ATT_TILE);
                const uint32_t idesc_s = make_idesc_bf16(ATT_TILE, nk, false);
                const uint64_t kd = k_desc0 + static_cast<uint64_t>((j % STAGES) * (ATT_TILE_BYTES >> 4));
#pragma unroll
                for (int k = 0; k < ATT_HD / 16; ++k) umma_ss(tS, q_desc + 2 * k, kd + 2 * k, idesc_s, k != 0 ? 1u : 0u);
                umma_commit(s_full);
            };
            auto issue_pv = [&](int j) {                  // O += P V_j
                const int nk = min(ATT_TILE, p.tokens - j * ATT_TILE);
                const uint64_t vd = v_desc0 + static_cast<uint64_t>((j % STAGES) * (ATT_TILE_BYTES >> 4));
                if (nk == ATT_TILE) {
#pragma unroll
                    for (int ks = 0; ks < ATT_TILE / 16; ++ks) umma_ts(tO, tP + ks * 8, vd + 128 * ks, idesc_pv, (j | ks) != 0 ? 1u : 0u);
                } else {
                    for (int ks = 0; ks < nk / 16; ++ks) umma_ts(tO, tP + ks * 8, vd + 128 * ks, idesc_pv, (j | ks) != 0 ? 1u : 0u);
                }
                umma_commit(pv_done);
                umma_commit(&kv_empty[j % STAGES]);
            };
            mbar_wait(q_full, 0);
            mbar_wait(&kv_full[0], 0);
            tc_fence_after();
            issue_s(0);
            for (int j = 0; j < nkv; ++j) {
                if (j + 1 < nkv) {
                    mbar_wait(&kv_full[(j + 1) % STAGES], ((j + 1) / STAGES) & 1);
                    mbar_wait(s_free, j & 1);
                    tc_fence_after();
                    issue_s(j + 1);
                }
                mbar_wait(p_ready, j & 1);
                tc_fence_after();
                issue_pv(j);
            }
        }
    } else {
        const int quarter = warp & 3, half = warp >> 2;
        const int row = quarter * 32 + lane;
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
        const uint32_t tS = t_lane + Cfg::S_COL + half * 64;
        const uint32_t tO = t_lane + Cfg::O_COL + half * 32;
        const uint32_t tP = t_lane + Cfg::P_COL + half * 32;
        const float c = p.scale_log2;
        const uint64_t c2 = pack_f32x2(c, c);
        const uint64_t zero2 = pack_f32x2(p.zero, p.zero);
        float m_ref = -INFINITY;
        uint64_t l2 = 0ull, l2b = 0ull;
        for (int j = 0; j < nkv; ++j) {
            const int nk = min(ATT_TILE, p.tokens - j * ATT_TILE) - half * 64;    // valid columns of this thread's half (may be <= 0)
            mbar_wait(s_full, j & 1);
            tc_fence_after();
            uint32_t r[64];
            if (nk >= 64) {
                tmem_ld32(tS, r);
                tmem_ld32(tS + 32, r + 32);
            } else {
#pragma unroll
                for (int col = 0; col < 64; col += 16) {
                    if (col < nk) {
                        tmem_ld16(tS + col, r + col);
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i) r[col + i] = 0xff800000u;   // -inf: exp2 -> 0, never read by P.V
                    }
                }
            }
            tmem_wait_ld();
            float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
            for (int i = 0; i < 64; i += 8) {
                m0 = fmax3(m0, __uint_as_float(r[i]), __uint_as_float(r[i + 1]));
                m1 = fmax3(m1, __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
                m2 = fmax3(m2, __uint_as_float(r[i + 4]), __uint_as_float(r[i + 5]));
                m3 = fmax3(m3, __uint_as_float(r[i + 6]), __uint_as_float(r[i + 7]));
            }
            // the pair's common view of the tile's row max: both halves rounded UP to bf16 (so the reference never lies
            // below a score by more than the lazy-rescale slack), exchanged through shared memory
            const __nv_bfloat16 m_loc_b = __float2bfloat16_ru(fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));
            float mt;
            if (DBG & 2) {                                  // diagnostic: no exchange (numerically wrong, timing only)
                mt = __bfloat162float(m_loc_b);
            } else {
                xch[half * 128 + row] = m_loc_b;
                pair_bar_sync(quarter);
                mt = fmaxf(__bfloat162float(m_loc_b), __bfloat162float(xch[(half ^ 1) * 128 + row]));
            }
            // S(j) is released only now: s_free(j) completing then also means that all eight warps have read this tile's
            // exchange slots, so the writes of tile j + 1 (which follow s_full(j + 1)) cannot overtake a read of tile j
            tc_fence_before();
            __syncwarp();
            if (elect_one()) mbar_arrive(s_free);
            if (j == 0) m_ref = mt;
            const bool need = (mt - m_ref) * c > ATT_RESCALE_LOG2;
            bool pv_waited = false;
            if (__any_sync(0xffffffffu, need)) {
                if (j > 0) { mbar_wait(pv_done, (j - 1) & 1); tc_fence_after(); pv_waited = true; }   // O is quiescent
                const float m_new = fmaxf(m_ref, mt);
                const float sc = ex2_approx((m_ref - m_new) * c);
                l2 = ffma2(l2, pack_f32x2(sc, sc), 0ull);
                l2b = ffma2(l2b, pack_f32x2(sc, sc), 0ull);
#pragma unroll
                for (int cidx = 0; cidx < 32; cidx += 16) {
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
            exp_chunk<DBG>(r, pk, c2, nmc2, zero2, l2, l2b);
            if (j > 0 && !pv_waited) { mbar_wait(pv_done, (j - 1) & 1); tc_fence_after(); }   // P(j-1) . V has retired
            tmem_st16(tP, pk);
            exp_chunk<DBG>(r + 32, pk, c2, nmc2, zero2, l2, l2b);
            tmem_st16(tP + 16, pk);
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (elect_one()) mbar_arrive(p_ready);
        }
        mbar_wait(pv_done, (nkv - 1) & 1);
        tc_fence_after();
        // the two partial row sums meet in the (now idle) first K stage
        float la, lb;
        unpack_f32x2(fadd2(l2, l2b), la, lb);
        float* lx = reinterpret_cast<float*>(sK);
        lx[half * 128 + row] = la + lb;
        pair_bar_sync(quarter);
        const float inv = 1.0f / (la + lb + lx[(half ^ 1) * 128 + row]);
        const int q = q0 + row;
        uint4 packed[4];
#pragma unroll
        for (int cidx = 0; cidx < 32; cidx += 16) {
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
            uint4* dst = reinterpret_cast<uint4*>(p.out + (static_cast<long long>(copy) * p.tokens + q) * hidden + head * ATT_HD + half * 32);
#pragma unroll
            for (int i = 0; i < 4; ++i) dst[i] = packed[i];
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == Cfg::W_MMA) tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
}

}  // namespace b200x

using namespace b200x;
namespace b200x { int attention_pingpong(const void* d_qkv, void* d_out, int copies, int tokens, int heads, int split, int reverse, long long* prof, int var, cudaStream_t s); }

// production configuration: one query tile per CTA (two CTAs per SM) with a quarter of the exponentials on the FMA pipe
// (variant bit 256); the other variants are diagnostics selected through the two b200x_debug_* setters below
static int g_attn_dbg = 256;
static int g_attn_nq = 1;

template <int DBG, int NQ = 2>
static int launch_attention(const CUtensorMap& tm, const AttnParams& p, dim3 grid, cudaStream_t s) {
    static bool configured = false;
    if (!configured) {
        B200X_CUDA_TRY(cudaFuncSetAttribute(attention_kernel<DBG, NQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttCfg<NQ>::SMEM));
        B200X_CUDA_TRY(cudaFuncSetAttribute(attention_kernel<DBG, NQ>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        configured = true;
    }
    attention_kernel<DBG, NQ><<<grid, AttCfg<NQ>::THREADS, AttCfg<NQ>::SMEM, s>>>(tm, p);
    B200X_CUDA_TRY(cudaGetLastError());
    return B200X_OK;
}

template <int DBG>
static int launch_attention_split(const CUtensorMap& tm, const AttnParams& p, dim3 grid, cudaStream_t s) {
    static bool configured = false;
    if (!configured) {
        B200X_CUDA_TRY(cudaFuncSetAttribute(attention_split_kernel<DBG>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttSplit::SMEM));
        B200X_CUDA_TRY(cudaFuncSetAttribute(attention_split_kernel<DBG>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        configured = true;
    }
    attention_split_kernel<DBG><<<grid, AttSplit::THREADS, AttSplit::SMEM, s>>>(tm, p);
    B200X_CUDA_TRY(cudaGetLastError());
    return B200X_OK;
}

template <bool POLY>
static int launch_attention_k64(const CUtensorMap& tmQ, const CUtensorMap& tmKV, const AttnParams& p, dim3 grid, cudaStream_t s) {
    static bool configured = false;
    if (!configured) {
        B200X_CUDA_TRY(cudaFuncSetAttribute(attention_k64_kernel<POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttK64::SMEM));
        B200X_CUDA_TRY(cudaFuncSetAttribute(attention_k64_kernel<POLY>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        configured = true;
    }
    attention_k64_kernel<POLY><<<grid, AttK64::THREADS, AttK64::SMEM, s>>>(tmQ, tmKV, p);
    B200X_CUDA_TRY(cudaGetLastError());
    return B200X_OK;
}

template <bool POLY>
static int launch_attention_fwd(const CUtensorMap& tm, const AttnParams& p, dim3 grid, cudaStream_t s) {
    static bool configured = false;
    if (!configured) {
        B200X_CUDA_TRY(cudaFuncSetAttribute(attention_fwd_kernel<POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttCfg<1>::SMEM));
        B200X_CUDA_TRY(cudaFuncSetAttribute(attention_fwd_kernel<POLY>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        configured = true;
    }
    attention_fwd_kernel<POLY><<<grid, AttCfg<1>::THREADS, AttCfg<1>::SMEM, s>>>(tm, p);
    B200X_CUDA_TRY(cudaGetLastError());
    return B200X_OK;
}

// diagnostic only (not part of the public header): select a stripped-down variant of the kernel for bottleneck analysis
static long long* g_attn_prof = nullptr;
extern "C" void b200x_debug_attention_variant(int v) { g_attn_dbg = v; }
extern "C" void b200x_debug_attention_tiles_per_cta(int nq) { g_attn_nq = nq; }
// diagnostic variant 32 writes per-CTA cycle counters ([cta][10][4] long long) to this device buffer
extern "C" void b200x_debug_attention_profile(void* d_buf) { g_attn_prof = static_cast<long long*>(d_buf); }

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
    AttnParams p{tokens, heads, reinterpret_cast<__nv_bfloat16*>(d_out), 0.125f * 1.4426950408889634f, g_attn_prof, 0.0f,
                 (g_attn_nq == 1 || g_attn_nq == 2 || g_attn_nq == 4) ? g_traverse_reverse : 0};
    dim3 grid(ceil_div(tokens, (g_attn_nq == 2 ? 2 : 1) * ATT_TILE), heads, copies);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (g_attn_nq == 5 || g_attn_nq == 6)                  // ping-pong kernel (attention_pingpong.cu), 4 or 8 softmax warps
        return attention_pingpong(d_qkv, d_out, copies, tokens, heads, g_attn_nq - 4, g_traverse_reverse, g_attn_prof, g_attn_dbg == 256 ? 0 : g_attn_dbg, s);
    if (g_attn_nq == 4) {                                  // 64-key tiles, P in place, three CTAs per SM
        CUtensorMap tmKV;
        const uint32_t box_kv[3] = {ATT_HD, AttK64::KV, 1};
        B200X_TRY(make_tmap_bf16(&tmKV, d_qkv, 3, dims, strides, box_kv));
        grid.x = ceil_div(tokens, ATT_TILE);
        switch (g_attn_dbg) {
            case 0: return launch_attention_k64<false>(tm, tmKV, p, grid, s);
            case 256: return launch_attention_k64<true>(tm, tmKV, p, grid, s);
            default: return set_error(B200X_ERR_INVALID, "attention: unknown variant %d for the 64-key kernel", g_attn_dbg);
        }
    }
    if (g_attn_nq == 3) {                                  // production kernel (software-pipelined softmax loop)
        grid.x = ceil_div(tokens, ATT_TILE);
        switch (g_attn_dbg) {
            case 0: return launch_attention_fwd<false>(tm, p, grid, s);
            case 256: return launch_attention_fwd<true>(tm, p, grid, s);
            default: return set_error(B200X_ERR_INVALID, "attention: unknown variant %d for the production kernel", g_attn_dbg);
        }
    }
    if (g_attn_nq == 0) {                                  // split-row kernel: one tile per CTA, two threads per query row
        grid.x = ceil_div(tokens, ATT_TILE);
        switch (g_attn_dbg) {
            case 0: return launch_attention_split<0>(tm, p, grid, s);
            case 256: return launch_attention_split<256>(tm, p, grid, s);
            case 258: return launch_attention_split<258>(tm, p, grid, s);
            case 16386: return launch_attention_split<16386>(tm, p, grid, s);
            default: return set_error(B200X_ERR_INVALID, "attention: unknown diagnostic variant %d for the split-row kernel", g_attn_dbg);
        }
    }
    if (g_attn_nq == 1) {
        switch (g_attn_dbg) {
            case 0: return launch_attention<0, 1>(tm, p, grid, s);
            case 32: return launch_attention<32, 1>(tm, p, grid, s);
            case 256: return launch_attention<256, 1>(tm, p, grid, s);
            case 33024: return launch_attention<33024, 1>(tm, p, grid, s);
            case 8192: return launch_attention<8192, 1>(tm, p, grid, s);
            case 16384: return launch_attention<16384, 1>(tm, p, grid, s);
            case 4096: return launch_attention<4096, 1>(tm, p, grid, s);
            case 4352: return launch_attention<4352, 1>(tm, p, grid, s);
            default: return set_error(B200X_ERR_INVALID, "attention: unknown diagnostic variant %d for one tile per CTA", g_attn_dbg);
        }
    }
    switch (g_attn_dbg) {
        case 0: return launch_attention<0>(tm, p, grid, s);
        case 4: return launch_attention<4>(tm, p, grid, s);
        case 12: return launch_attention<12>(tm, p, grid, s);
        case 20: return launch_attention<20>(tm, p, grid, s);
        case 28: return launch_attention<28>(tm, p, grid, s);
        case 1: return launch_attention<1>(tm, p, grid, s);
        case 32: return launch_attention<32>(tm, p, grid, s);
        case 96: return launch_attention<96>(tm, p, grid, s);
        case 24: return launch_attention<24>(tm, p, grid, s);
        case 88: return launch_attention<88>(tm, p, grid, s);
        case 120: return launch_attention<120>(tm, p, grid, s);
        case 512: return launch_attention<512>(tm, p, grid, s);
        case 1024: return launch_attention<1024>(tm, p, grid, s);
        case 65536: return launch_attention<65536>(tm, p, grid, s);
        case 65792: return launch_attention<65792>(tm, p, grid, s);
        case 69888: return launch_attention<69888>(tm, p, grid, s);
        case 4096: return launch_attention<4096>(tm, p, grid, s);
        case 4352: return launch_attention<4352>(tm, p, grid, s);
        case 3072: return launch_attention<3072>(tm, p, grid, s);
        case 576: return launch_attention<576>(tm, p, grid, s);
        case 608: return launch_attention<608>(tm, p, grid, s);
        case 64: return launch_attention<64>(tm, p, grid, s);
        case 128: return launch_attention<128>(tm, p, grid, s);
        case 256: return launch_attention<256>(tm, p, grid, s);
        case 320: return launch_attention<320>(tm, p, grid, s);
        case 384: return launch_attention<384>(tm, p, grid, s);
        default: return set_error(B200X_ERR_INVALID, "attention: unknown diagnostic variant %d", g_attn_dbg);
    }
}
