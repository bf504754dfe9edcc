// Ping-pong fused self-attention for the SpecTTTra encoder on sm_100a (head_dim 64, bf16 in, fp32 accumulate).
//
// Replaces F.scaled_dot_product_attention of the third-party `sonics` encoder block (reference call site:
// src/sonics_api.py:259-271 -> HFAudioClassifier forward).
//
// One persistent CTA per SM works on TWO independent query tiles ("streams" A and B, 128 query rows each, any
// (copy, head, tile) items) and the SAME softmax warps alternate between them:
//     softmax pass A(g) | tensor pipe: S_B(g) .. PV_B(g-1)        softmax pass B(g) | tensor pipe: S_A(g+1), PV_A(g)
// so the MUFU / FMA work of one tile always runs against the MMAs of the other one, with no second group of softmax
// warps competing for the same MUFU (the round-1 kernel ran two free-running CTAs per SM whose exponential phases
// collided at random: its four lane-quarter warps finished up to 1600 cycles apart and every rendezvous waited for the
// slowest one).  The softmax is a SINGLE pass per key tile: the exponentials are taken against the running reference
// maximum while the tile's own maximum is tracked on the side; only if some row exceeds the reference by more than 2^8
// (rare after the first key tile) the pass is redone with the new maximum and O / l are rescaled.  The first key tile of
// an item takes one extra max-only pass.
//   warps [0, 4 SPLIT) : softmax.  TMEM lane == query row; with SPLIT = 2 two warps share a lane quarter and take half
//                        of the key columns each (row statistics exchanged through shared memory, off the critical path)
//   warp 4 SPLIT       : TMA producer (Q per item; K / V through a 3-stage mbarrier ring per stream)
//   warp 4 SPLIT + 1   : TMEM allocator + single-thread tcgen05.mma issuer for both streams
// TMEM columns: S_A [0,128) S_B [128,256) O_A [256,320) O_B [320,384) P_A [384,448) P_B [448,512).
#include "common.h"
#include "ptx.cuh"

namespace b200x {

constexpr int PP_TILE = 128;
constexpr int PP_HD = 64;
constexpr int PP_TILE_BYTES = PP_TILE * PP_HD * 2;       // 16 KB
constexpr int PP_STAGES = 3;
constexpr float PP_RESCALE_LOG2 = 8.0f;

template <int SPLIT> struct PPCfg {
    static constexpr int SM_WARPS = 4 * SPLIT;
    static constexpr int W_TMA = SM_WARPS, W_MMA = SM_WARPS + 1;
    static constexpr int THREADS = 32 * (SM_WARPS + 2);
    static constexpr int TILES = 2 + 2 * 2 * PP_STAGES;                 // Q_A Q_B + (K, V) x stages x streams
    static constexpr int XCH_OFFSET = TILES * PP_TILE_BYTES;            // float xch[2][2][128]
    static constexpr int BAR_OFFSET = XCH_OFFSET + 2048;
    static constexpr int SMEM = BAR_OFFSET + 256;
    static_assert(SMEM <= 232448, "shared memory budget exceeded");
};

struct PPParams {
    int tokens, heads, copies;
    int n_qt, n_items;
    __nv_bfloat16* out;      // [copies * tokens, heads * 64]
    float scale_log2;        // (1/sqrt(64)) * log2(e)
    float zero;              // always 0.0f: an operand ptxas cannot fold (see exp32)
    int reverse;             // walk the item list from its end (L2 reuse of the QKV GEMM's last output)
    long long* prof;         // PROF instantiation only: [cta][16] cycle counters of softmax warp 0 and of the MMA issuer
};

// 32 scores -> p = 2^(s*c - m*c) -> 16 packed bf16 pairs; row sums into two packed accumulators.  A quarter of the pairs
// take the FMA-pipe polynomial (exp2_poly2) instead of the MUFU.  The two 16-score halves take their addend through a
// data dependence on the running sums (value unchanged) so that ptxas keeps the MUFU runs short and interleaved.
template <int VAR>
__device__ __forceinline__ void exp32(const uint32_t* r, uint32_t (&pk)[16], uint64_t c2, uint64_t nmc2, uint64_t zero2,
                                      uint64_t& acc_a, uint64_t& acc_b) {
    const uint64_t link_a = ffma2(acc_a, zero2, nmc2);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float x0, x1, p0, p1;
        unpack_f32x2(ffma2(pack_f32x2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), c2, link_a), x0, x1);
        if ((VAR & 16) ? (i & 1) : (i & 3) == 3) exp2_poly2(x0, x1, p0, p1);
        else { p0 = (VAR & 1) ? x0 : ex2_approx(x0); p1 = (VAR & 1) ? x1 : ex2_approx(x1); }
        acc_a = fadd2(acc_a, pack_f32x2(p0, p1));
        pk[i] = pack_bf16(p0, p1);
    }
    const uint64_t link_b = ffma2(acc_b, zero2, nmc2);
#pragma unroll
    for (int i = 8; i < 16; ++i) {
        float x0, x1, p0, p1;
        unpack_f32x2(ffma2(pack_f32x2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), c2, link_b), x0, x1);
        if ((VAR & 16) ? (i & 1) : (i & 3) == 3) exp2_poly2(x0, x1, p0, p1);
        else { p0 = (VAR & 1) ? x0 : ex2_approx(x0); p1 = (VAR & 1) ? x1 : ex2_approx(x1); }
        acc_b = fadd2(acc_b, pack_f32x2(p0, p1));
        pk[i] = pack_bf16(p0, p1);
    }
}

__device__ __forceinline__ float max32(const uint32_t* r, float m) {
    float m0 = m, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
        m0 = fmax3(m0, __uint_as_float(r[i]), __uint_as_float(r[i + 1]));
        m1 = fmax3(m1, __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
        m2 = fmax3(m2, __uint_as_float(r[i + 4]), __uint_as_float(r[i + 5]));
        m3 = fmax3(m3, __uint_as_float(r[i + 6]), __uint_as_float(r[i + 7]));
    }
    return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
}

// columns [valid, 32) of a chunk lie past the last key: they read as -inf (valid is a multiple of 16)
__device__ __forceinline__ void mask32(uint32_t* r, int valid) {
    if (valid < 32) {
#pragma unroll
        for (int i = 16; i < 32; ++i) r[i] = 0xff800000u;
        if (valid < 16) {
#pragma unroll
            for (int i = 0; i < 16; ++i) r[i] = 0xff800000u;
        }
    }
}

template <int SPLIT, bool PROF, int VAR>
__global__ void __launch_bounds__(PPCfg<SPLIT>::THREADS, 1)
attention_pingpong_kernel(const __grid_constant__ CUtensorMap tmQKV, PPParams p) {
    using Cfg = PPCfg<SPLIT>;
    constexpr int COLS = PP_TILE / SPLIT;                 // score columns per softmax thread
    constexpr int NCH = COLS / 32;                        // 32-column chunks per thread
    constexpr int OCOLS = PP_HD / SPLIT;                  // output columns per thread
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sQ = smem;                                   // [2] tiles
    uint8_t* sKV = smem + 2 * PP_TILE_BYTES;              // [stream][stage]{K, V}
    float* xch = reinterpret_cast<float*>(smem + Cfg::XCH_OFFSET);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFFSET);
    uint64_t* q_full = bars;                              // [2]
    uint64_t* q_empty = q_full + 2;                       // [2]
    uint64_t* kv_full = q_empty + 2;                      // [2][STAGES]
    uint64_t* kv_empty = kv_full + 2 * PP_STAGES;         // [2][STAGES]
    uint64_t* s_full = kv_empty + 2 * PP_STAGES;          // [2]
    uint64_t* p_ready = s_full + 2;                       // [2]
    uint64_t* pv_done = p_ready + 2;                      // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int nkv = (p.tokens + PP_TILE - 1) / PP_TILE;
    const int hidden = p.heads * PP_HD;
    // stream X of CTA b owns items 2 b + X + 2 G k (k = 0, 1, ...): at any time the grid works on a contiguous item range
    const int G2 = 2 * gridDim.x;
    int n_k[2], steps[2];
#pragma unroll
    for (int X = 0; X < 2; ++X) {
        const int first = 2 * blockIdx.x + X;
        n_k[X] = first < p.n_items ? (p.n_items - first + G2 - 1) / G2 : 0;
        steps[X] = n_k[X] * nkv;
    }
    const int max_k = max(n_k[0], n_k[1]);
    const int max_steps = max_k * nkv;
    auto item_coords = [&](int X, int k, int& copy, int& head, int& qt) {
        int it = 2 * blockIdx.x + X + G2 * k;
        if (p.reverse) it = p.n_items - 1 - it;
        qt = it % p.n_qt;
        head = (it / p.n_qt) % p.heads;
        copy = it / (p.n_qt * p.heads);
    };

    if (warp == Cfg::W_TMA && elect_one()) {
        tma_prefetch_desc(&tmQKV);
        for (int X = 0; X < 2; ++X) {
            mbar_init(&q_full[X], 1);
            mbar_init(&q_empty[X], 1);
            for (int s = 0; s < PP_STAGES; ++s) {
                mbar_init(&kv_full[X * PP_STAGES + s], 1);
                mbar_init(&kv_empty[X * PP_STAGES + s], 1);
            }
            mbar_init(&s_full[X], 1);
            mbar_init(&p_ready[X], Cfg::SM_WARPS);
            mbar_init(&pv_done[X], 1);
        }
        fence_barrier_init();
    }
    if (warp == Cfg::W_MMA) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == Cfg::W_TMA) {
        // ------------------------------------------------------------------ TMA producer
        if (elect_one()) {
            int g[2] = {0, 0};
            for (int k = 0; k < max_k; ++k) {
                int copy[2] = {0, 0}, head[2] = {0, 0}, qt[2] = {0, 0};
#pragma unroll
                for (int X = 0; X < 2; ++X) {
                    if (k >= n_k[X]) continue;
                    item_coords(X, k, copy[X], head[X], qt[X]);
                    mbar_wait(&q_empty[X], (k & 1) ^ 1);             // the S MMAs of item k - 1 have read Q
                    mbar_expect_tx(&q_full[X], PP_TILE_BYTES);
                    tma_load_3d(sQ + X * PP_TILE_BYTES, &tmQKV, &q_full[X], head[X] * PP_HD, qt[X] * PP_TILE, copy[X]);
                }
                for (int j = 0; j < nkv; ++j) {
#pragma unroll
                    for (int X = 0; X < 2; ++X) {
                        if (k >= n_k[X]) continue;
                        const int st = g[X] % PP_STAGES;
                        const uint32_t ph = (g[X] / PP_STAGES) & 1;
                        ++g[X];
                        uint64_t* full = &kv_full[X * PP_STAGES + st];
                        uint8_t* dst = sKV + (X * PP_STAGES + st) * 2 * PP_TILE_BYTES;
                        mbar_wait(&kv_empty[X * PP_STAGES + st], ph ^ 1);
                        mbar_expect_tx(full, 2 * PP_TILE_BYTES);
                        tma_load_3d(dst, &tmQKV, full, hidden + head[X] * PP_HD, j * PP_TILE, copy[X]);
                        tma_load_3d(dst + PP_TILE_BYTES, &tmQKV, full, 2 * hidden + head[X] * PP_HD, j * PP_TILE, copy[X]);
                    }
                }
            }
        }
    } else if (warp == Cfg::W_MMA) {
        // ------------------------------------------------------------------ MMA issuer (both streams)
        if (elect_one()) {
            constexpr uint32_t idesc_pv = make_idesc_bf16(PP_TILE, PP_HD, true);
            long long pc_p = 0, pc_kv = 0, pc_t = 0;
            const long long pc_start = PROF ? clock64() : 0;
            auto issue_s = [&](int X, int g) {            // S_X = Q_X K^T for step g of stream X
                const int k = g / nkv, j = g - k * nkv;
                if (PROF) pc_t = clock64();
                if (j == 0) mbar_wait(&q_full[X], k & 1);
                const int st = g % PP_STAGES;
                mbar_wait(&kv_full[X * PP_STAGES + st], (g / PP_STAGES) & 1);
                tc_fence_after();
                if (PROF) pc_kv += clock64() - pc_t;
                const int nk = min(PP_TILE, p.tokens - j * PP_TILE);
                const uint32_t idesc_s = make_idesc_bf16(PP_TILE, nk, false);
                const uint64_t qd = make_smem_desc_sw128(smem_u32(sQ + X * PP_TILE_BYTES), 16, 1024);
                const uint64_t kd = make_smem_desc_sw128(smem_u32(sKV + (X * PP_STAGES + st) * 2 * PP_TILE_BYTES), 16, 1024);
                const uint32_t tS = tmem_base + X * PP_TILE;
#pragma unroll
                for (int kk = 0; kk < PP_HD / 16; ++kk) umma_ss(tS, qd + 2 * kk, kd + 2 * kk, idesc_s, kk != 0 ? 1u : 0u);
                umma_commit(&s_full[X]);
                if (j == nkv - 1) umma_commit(&q_empty[X]);           // last use of this item's Q
            };
            auto issue_pv = [&](int X, int g) {           // O_X (+)= P_X V
                const int k = g / nkv, j = g - k * nkv;
                const int st = g % PP_STAGES;
                const int nk = min(PP_TILE, p.tokens - j * PP_TILE);
                const uint64_t vd = make_smem_desc_sw128(smem_u32(sKV + (X * PP_STAGES + st) * 2 * PP_TILE_BYTES + PP_TILE_BYTES), 16384, 1024);
                const uint32_t tO = tmem_base + 256 + X * PP_HD, tP = tmem_base + 384 + X * PP_HD;
                if (nk == PP_TILE) {
#pragma unroll
                    for (int ks = 0; ks < PP_TILE / 16; ++ks) umma_ts(tO, tP + ks * 8, vd + 128 * ks, idesc_pv, (j | ks) != 0 ? 1u : 0u);
                } else {
                    for (int ks = 0; ks < nk / 16; ++ks) umma_ts(tO, tP + ks * 8, vd + 128 * ks, idesc_pv, (j | ks) != 0 ? 1u : 0u);
                }
                umma_commit(&pv_done[X]);
                umma_commit(&kv_empty[X * PP_STAGES + st]);
            };
#pragma unroll
            for (int X = 0; X < 2; ++X)
                if (steps[X] > 0) issue_s(X, 0);
            for (int g = 0; g < max_steps; ++g) {
#pragma unroll
                for (int X = 0; X < 2; ++X) {
                    if (g >= steps[X]) continue;
                    if (PROF) pc_t = clock64();
                    mbar_wait(&p_ready[X], g & 1);        // pass (X, g) is over: S_X is free, P_X is in TMEM
                    tc_fence_after();
                    if (PROF) pc_p += clock64() - pc_t;
                    if (g + 1 < steps[X]) issue_s(X, g + 1);
                    issue_pv(X, g);
                }
            }
            if (PROF) {
                long long* o = p.prof + blockIdx.x * 16 + 8;
                o[0] = pc_p; o[1] = pc_kv; o[2] = clock64() - pc_start; o[3] = max_steps;
            }
        }
    } else {
        // ------------------------------------------------------------------ softmax warps
        const int quarter = warp & 3;
        const int half = SPLIT == 2 ? (warp >> 2) : 0;
        const int row = quarter * 32 + lane;
        const int col0 = half * COLS;
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
        const float c = p.scale_log2;
        const uint64_t c2 = pack_f32x2(c, c);
        const uint64_t zero2 = pack_f32x2(p.zero, p.zero);
        int xn = 0;                                       // exchange counter (SPLIT == 2): buffer parity
        // value of the partner thread (same row, other column half); one named barrier per exchange, alternating buffers
        auto exchange = [&](float v) -> float {
            if (SPLIT == 1) return v;
            float* buf = xch + (xn & 1) * 256;
            ++xn;
            buf[half * 128 + row] = v;
            named_bar_sync(1 + quarter, 64);
            return buf[(half ^ 1) * 128 + row];
        };
        auto epilogue = [&](int X, int k, int g_last, uint64_t sum_a, uint64_t sum_b) {   // O_X / l -> bf16 rows of item k
            mbar_wait(&pv_done[X], g_last & 1);
            tc_fence_after();
            int copy, head, qt;
            item_coords(X, k, copy, head, qt);
            float a0, a1;
            unpack_f32x2(fadd2(sum_a, sum_b), a0, a1);
            float l = a0 + a1;
            if (SPLIT == 2) l += exchange(l);
            const float inv = 1.0f / l;
            const uint32_t tO = t_lane + 256 + X * PP_HD + half * OCOLS;
            uint4 packed[OCOLS / 8];
#pragma unroll
            for (int cidx = 0; cidx < OCOLS; cidx += 16) {
                uint32_t o[16];
                tmem_ld16(tO + cidx, o);
                tmem_wait_ld();
                uint32_t w[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) w[i] = pack_bf16(__uint_as_float(o[2 * i]) * inv, __uint_as_float(o[2 * i + 1]) * inv);
                packed[cidx / 8] = make_uint4(w[0], w[1], w[2], w[3]);
                packed[cidx / 8 + 1] = make_uint4(w[4], w[5], w[6], w[7]);
            }
            const int q = qt * PP_TILE + row;
            if (q < p.tokens) {
                uint4* dst = reinterpret_cast<uint4*>(p.out + (static_cast<long long>(copy) * p.tokens + q) * hidden + head * PP_HD + half * OCOLS);
#pragma unroll
                for (int i = 0; i < OCOLS / 8; ++i) dst[i] = packed[i];
            }
        };

        // ONE copy of the pass code serves both streams (X is a run-time value, the per-stream state is swapped between two
        // register sets) and the rare redo: the loop body must stay well inside the 32 KB instruction cache - with one
        // softmax warp per scheduler nothing hides an instruction fetch from L2.
        float m_cur = -INFINITY, m_oth = -INFINITY;
        uint64_t la_cur = 0ull, lb_cur = 0ull, la_oth = 0ull, lb_oth = 0ull;
        long long pc_s = 0, pc_first = 0, pc_chunks = 0, pc_pv = 0, pc_tail = 0, pc_t = 0, pc_u = 0, pc_n = 0;
        const long long pc_start = PROF ? clock64() : 0;
        const int steps0 = steps[0], steps1 = steps[1];
#pragma unroll 1
        for (int hstep = 0; hstep < 2 * max_steps; ++hstep) {
            const int X = hstep & 1, g = hstep >> 1;
            if (g < (X ? steps1 : steps0)) {
                const int k = g / nkv, j = g - k * nkv;
                if (j == 0 && k > 0) epilogue(X, k - 1, g - 1, la_cur, lb_cur);
                const int nk = min(PP_TILE, p.tokens - j * PP_TILE);
                const int vc = min(max(nk - col0, 0), COLS);          // my valid columns (multiple of 16)
                const uint32_t tS = t_lane + X * PP_TILE + col0;
                const uint32_t tP = t_lane + 384 + X * PP_HD + col0 / 2;
                if (PROF) pc_t = clock64();
                mbar_wait(&s_full[X], g & 1);
                tc_fence_after();
                if (PROF) { const long long t = clock64(); pc_s += t - pc_t; pc_t = t; ++pc_n; }
                uint32_t r[2][32];
                if (j == 0) {
                    // first key tile of an item: the reference maximum is the tile's true row maximum
                    float mt = -INFINITY;
#pragma unroll
                    for (int ch = 0; ch < NCH; ++ch) {
                        if (ch * 32 < vc) {
                            tmem_ld32(tS + ch * 32, r[0]);
                            tmem_wait_ld();
                            mask32(r[0], vc - ch * 32);
                            mt = max32(r[0], mt);
                        }
                    }
                    if (SPLIT == 2) mt = fmaxf(mt, exchange(mt));
                    m_cur = mt;
                    la_cur = 0ull;
                    lb_cur = 0ull;
                }
                if (PROF) { const long long t = clock64(); pc_first += t - pc_t; pc_t = t; }
                const uint64_t la_save = la_cur, lb_save = lb_cur;
#pragma unroll 1
                for (int attempt = 0; attempt < 2; ++attempt) {
                    float mt = -INFINITY;
                    const float mc = m_cur * c;
                    const uint64_t nmc2 = pack_f32x2(-mc, -mc);
                    if (vc > 0) tmem_ld32(tS, r[0]);
#pragma unroll
                    for (int ch = 0; ch < NCH; ++ch) {
                        if (ch * 32 < vc) {
                            tmem_wait_ld();
                            if (!(VAR & 8) && ch + 1 < NCH && (ch + 1) * 32 < vc) tmem_ld32(tS + (ch + 1) * 32, r[(ch + 1) & 1]);
                            uint32_t* rc = r[(VAR & 8) ? 0 : (ch & 1)];
                            mask32(rc, vc - ch * 32);
                            if (!(VAR & 4)) mt = max32(rc, mt);
                            uint32_t pk[16];
                            exp32<VAR>(rc, pk, c2, nmc2, zero2, la_cur, lb_cur);
                            if (ch == 0 && j > 0 && attempt == 0) {                  // P_X(g-1) . V retired
                                if (PROF) pc_u = clock64();
                                mbar_wait(&pv_done[X], (g - 1) & 1);
                                tc_fence_after();
                                if (PROF) pc_pv += clock64() - pc_u;
                            }
                            if (!(VAR & 2)) tmem_st16(tP + ch * 16, pk); else asm volatile("" :: "r"(pk[0] ^ pk[5] ^ pk[9] ^ pk[15]));
                        }
                    }
                    if (j == 0 || attempt == 1) break;
                    if (vc == 0) { mbar_wait(&pv_done[X], (g - 1) & 1); tc_fence_after(); }
                    if (SPLIT == 2) mt = fmaxf(mt, exchange(mt));
                    const bool need = (mt - m_cur) * c > PP_RESCALE_LOG2;
                    if (!__any_sync(0xffffffffu, need)) break;
                    // rare: redo the pass against the new maximum; O and l are rescaled (P.V of this step is not issued yet)
                    const float m_new = fmaxf(m_cur, mt);
                    const float sc = ex2_approx((m_cur - m_new) * c);
                    tmem_wait_st();
                    const uint32_t tO = t_lane + 256 + X * PP_HD + half * OCOLS;
#pragma unroll
                    for (int cidx = 0; cidx < OCOLS; cidx += 16) {
                        uint32_t o[16];
                        tmem_ld16(tO + cidx, o);
                        tmem_wait_ld();
#pragma unroll
                        for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * sc);
                        tmem_st16(tO + cidx, o);
                    }
                    la_cur = ffma2(la_save, pack_f32x2(sc, sc), 0ull);
                    lb_cur = ffma2(lb_save, pack_f32x2(sc, sc), 0ull);
                    m_cur = m_new;
                }
                if (PROF) { const long long t = clock64(); pc_chunks += t - pc_t; pc_t = t; }
                tmem_wait_st();
                tc_fence_before();
                __syncwarp();
                if (elect_one()) mbar_arrive(&p_ready[X]);
                if (PROF) pc_tail += clock64() - pc_t;
            }
            // the other stream's turn
            const float tm = m_cur; m_cur = m_oth; m_oth = tm;
            const uint64_t ta = la_cur; la_cur = la_oth; la_oth = ta;
            const uint64_t tb = lb_cur; lb_cur = lb_oth; lb_oth = tb;
        }
        // after 2 * max_steps half steps the "current" set belongs to stream 0 again
        if (steps0 > 0) epilogue(0, n_k[0] - 1, steps0 - 1, la_cur, lb_cur);
        if (steps1 > 0) epilogue(1, n_k[1] - 1, steps1 - 1, la_oth, lb_oth);
        if (PROF && warp == 0 && lane == 0) {
            long long* o = p.prof + blockIdx.x * 16;
            o[0] = pc_s; o[1] = pc_first; o[2] = pc_chunks; o[3] = pc_pv; o[4] = pc_tail; o[5] = clock64() - pc_start; o[6] = pc_n;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == Cfg::W_MMA) tmem_dealloc<512>(tmem_base);
}

template <int SPLIT, bool PROF, int VAR = 0>
static int launch_pingpong(const CUtensorMap& tm, const PPParams& p, int grid, cudaStream_t s) {
    using Cfg = PPCfg<SPLIT>;
    B200X_CUDA_TRY(cudaFuncSetAttribute(attention_pingpong_kernel<SPLIT, PROF, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attention_pingpong_kernel<SPLIT, PROF, VAR><<<grid, Cfg::THREADS, Cfg::SMEM, s>>>(tm, p);
    B200X_CUDA_TRY(cudaGetLastError());
    return B200X_OK;
}

int attention_pingpong(const void* d_qkv, void* d_out, int copies, int tokens, int heads, int split, int reverse, long long* prof,
                       int var, cudaStream_t s) {
    const int width = 3 * heads * PP_HD;
    CUtensorMap tm;
    const uint64_t dims[3] = {static_cast<uint64_t>(width), static_cast<uint64_t>(tokens), static_cast<uint64_t>(copies)};
    const uint64_t strides[2] = {static_cast<uint64_t>(width) * 2, static_cast<uint64_t>(width) * 2 * tokens};
    const uint32_t box[3] = {PP_HD, PP_TILE, 1};
    B200X_TRY(make_tmap_bf16(&tm, d_qkv, 3, dims, strides, box));
    PPParams p;
    p.tokens = tokens;
    p.heads = heads;
    p.copies = copies;
    p.n_qt = ceil_div(tokens, PP_TILE);
    p.n_items = copies * heads * p.n_qt;
    p.out = reinterpret_cast<__nv_bfloat16*>(d_out);
    p.scale_log2 = 0.125f * 1.4426950408889634f;
    p.zero = 0.0f;
    p.reverse = reverse;
    p.prof = prof;
    int dev = 0, sms = 0;
    B200X_CUDA_TRY(cudaGetDevice(&dev));
    B200X_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int grid = std::min(sms, ceil_div(p.n_items, 2));
    if (split == 1) {
        switch (var) {
            case 1: return launch_pingpong<1, false, 1>(tm, p, grid, s);
            case 2: return launch_pingpong<1, false, 2>(tm, p, grid, s);
            case 3: return launch_pingpong<1, false, 3>(tm, p, grid, s);
            case 4: return launch_pingpong<1, false, 4>(tm, p, grid, s);
            case 8: return launch_pingpong<1, false, 8>(tm, p, grid, s);
            case 15: return launch_pingpong<1, false, 15>(tm, p, grid, s);
            case 16: return launch_pingpong<1, false, 16>(tm, p, grid, s);
            default: break;
        }
    }
    if (prof != nullptr) return split == 2 ? launch_pingpong<2, true>(tm, p, grid, s) : launch_pingpong<1, true>(tm, p, grid, s);
    return split == 2 ? launch_pingpong<2, false>(tm, p, grid, s) : launch_pingpong<1, false>(tm, p, grid, s);
}

}  // namespace b200x
