// REJECTED VARIANT (measured, not part of libb200xai.so): LayerNorm fused into the CONSUMER projection (QKV / fc1) with an
// A-stationary CTA-pair GEMM.  Producer warps normalise fp32 rows into an L2-resident scratch tile, a TMA-A warp refills the
// stationary shared-memory A operand from it, and all column tiles of W sweep over the row tile.
// Bit-identical to layernorm + gemm (tests passed), but slower than the pair of kernels it replaces on B200
// (profiles/r02_g_gemm_ln_astationary.txt, 229 copies, same run):
//   layernorm 115 us + QKV gemm 269 us = 384 us   vs   fused 383-439 us;   fc1: 115 + 299 us   vs   591-697 us
//   without the LayerNorm work (debug=1) the A-stationary pipeline alone runs QKV in 252-265 us (1.05-1.11 PFLOP/s).
// In-kernel cycle counters: the eight producer warps need ~34 k cycles per 128-row tile (load 2.1 k + variance 1.5 k +
// normalise/store 2.2 k per 4-row batch, + 3.6 k for the gpu-scope fence before the TMA hand-over) against ~25 k for the
// MMA pipeline, and they take the registers (608 threads -> 96 per thread) the GELU epilogue needs (168).
// The production answer is the opposite fusion: LayerNorm as a TAIL of the residual GEMM that produces x
// (gemm2_resid_ln_kernel in csrc/gemm_tcgen05.cu) - those kernels are HBM-bound and have the issue slots to spare.
// This file is an excerpt (it needs the helpers of csrc/gemm_tcgen05.cu around it to compile).

// ------------------------------------------------------------------------------------------------ LayerNorm-fused CTA-pair GEMM
// out = act(LayerNorm(x) . W^T + bias), bf16: the LayerNorm that precedes the QKV and fc1 projections (pre-norm encoder block)
// runs INSIDE the GEMM that consumes it.  The separate pass read 484 MB and wrote 242 MB per call at 229 copies (9 % of the
// step) and the GEMM then read the 242 MB again - six times over through L2, once per column tile.
// A-STATIONARY: a CTA keeps the normalised 128 x K row tile (K <= 384: six 16 KB k-blocks, 128-byte swizzled) in shared
// memory and sweeps ALL column tiles of W over it, so the L2 -> SM traffic per row tile drops from 6 x (A + B) to A + 6 x B
// (the plain pair kernel is paced by exactly that traffic, see above) and x is read from HBM once.
//   warps 10-17 : LayerNorm producers, one warp per row, the arithmetic of layernorm_kernel bit for bit.  They run up to two
//                 row tiles AHEAD of the MMAs and write the normalised bf16 rows to a per-CTA, double-buffered scratch tile in
//                 global memory (2 x 96 KB per CTA, 28 MB per launch: rewritten every few microseconds, it lives in L2).
//   warp 18     : TMA-A: as soon as the MMAs of the previous row tile release a k-block (a_empty[kb], during its LAST column
//                 tile) the next row tile's k-block is fetched from the scratch tile - the refill overlaps the remaining MMAs.
//   warp 0      : TMA producer of the B (weight) half tiles, 3/4-stage ring, runs ahead across row-tile boundaries
//   warp 1      : (leader) MMA issuer: waits a_full[kb] on the first column tile of a row tile, commits a_empty[kb] to both
//                 CTAs on the last one
//   warps 2-9   : epilogue, unchanged (TMEM -> bias / GELU -> bf16 slab -> TMA store); the first column tile's "accumulator
//                 full" also proves that the whole A tile has landed, i.e. that its scratch buffer may be rewritten
// Keeping the producers off the MMA's critical path is the point of the scratch tile: a first version that normalised
// straight into the shared-memory operand had to wait for each k-block to be released before it could even fetch the rows
// that replace it (283 us against 78 + 37 us for the separate kernels, profiles/r02_f_gemm_ln_v1.txt).
constexpr int GEMM_LN_PRODUCER_WARPS = 8;
constexpr int GEMM_LN_W_PRODUCER = 2 + GEMM_EPI_WARPS;                       // first producer warp
constexpr int GEMM_LN_W_TMA_A = GEMM_LN_W_PRODUCER + GEMM_LN_PRODUCER_WARPS;
constexpr int GEMM_LN_THREADS = 32 * (GEMM_LN_W_TMA_A + 1);
constexpr int GEMM_LN_MAX_KB = 6;
constexpr int GEMM_LN_ROWS_PER_BATCH = 4;    // rows a producer warp has in flight (12 independent 16-byte loads per lane)

struct GemmLnParams {
    const float* x;          // [M][ldx] fp32 residual stream
    int ldx;
    const float* gamma;      // [K]
    const float* beta;       // [K]
    float eps;
    __nv_bfloat16* scratch;  // [gridDim.x][2][128][K] normalised row tiles (tmA maps it as [gridDim.x * 256][K])
    int debug;               // experiment switches (0 in production)
};

template <int BN>
struct GemmLnSmem {
    static constexpr int A_KB_BYTES = GEMM_BM * GEMM_BK * 2;                  // 16 KB per k-block
    static constexpr int A_BYTES = GEMM_LN_MAX_KB * A_KB_BYTES;               // 96 KB
    static constexpr int B_BYTES = (BN / 2) * GEMM_BK * 2;
    // the weight ring is as deep as shared memory allows (7-8 stages): with A stationary a stage is only the B half tile and
    // the ring must cover the L2 -> SM round trip (~3 k cycles measured, i.e. 8 k-steps of 384 MMA cycles); the epilogue
    // warps make do with one output slab each
    static constexpr int EPI_WARP_BYTES = GEMM_SLAB_BYTES;
    static constexpr int STAGES_FIT = (232448 - 1024 - 512 - A_BYTES - GEMM_EPI_WARPS * EPI_WARP_BYTES) / B_BYTES;
    static constexpr int STAGES = STAGES_FIT > 8 ? 8 : STAGES_FIT;
    static constexpr int B_OFFSET = A_BYTES;
    static constexpr int EPI_OFFSET = B_OFFSET + STAGES * B_BYTES;
    static constexpr int BAR_OFFSET = EPI_OFFSET + GEMM_EPI_WARPS * EPI_WARP_BYTES;
    static constexpr int TOTAL = BAR_OFFSET + 512 + 1024;
    static_assert(B_BYTES % 1024 == 0, "B half tile must keep 1024-byte alignment for the 128B swizzle");
    static_assert(STAGES >= 4 && TOTAL <= 232448, "shared memory budget exceeded");
};

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_LN_THREADS, 1)
gemm2_ln_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                        const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmCtail, GemmParams p,
                        GemmLnParams q) {
    using L = GemmLnSmem<BN>;
    constexpr uint32_t TMEM_COLS = (2 * BN <= 256) ? 256 : 512;
    constexpr int STAGES = L::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint64_t* a_full = tempty_bar + 2;                   // [6] k-block kb of the row tile has landed in BOTH CTAs (leader's copy)
    uint64_t* a_empty = a_full + GEMM_LN_MAX_KB;         // [6] per CTA: the row tile's last MMAs on k-block kb have retired
    uint64_t* sc_full = a_empty + GEMM_LN_MAX_KB;        // [2] per CTA: scratch buffer written by the producer warps
    uint64_t* sc_empty = sc_full + 2;                    // [2] per CTA: scratch buffer read completely by TMA-A
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sc_empty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int m_tiles = (p.M + 2 * GEMM_BM - 1) / (2 * GEMM_BM);
    const int n_tiles = (p.N + BN - 1) / BN;
    const int num_kb = (p.K + GEMM_BK - 1) / GEMM_BK;
    // the pairs walk the column tiles in rotated orders: in lockstep all 74 pairs would pull the SAME 24 KB weight tile out of
    // the same few L2 slices at the same time
    const int n_rot = (q.debug & 16) ? 0 : pair % n_tiles;

    if (warp == 0 && elect_one()) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmC);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tfull_bar[s], 1);
            mbar_init(&tempty_bar[s], 2 * GEMM_EPI_WARPS);
            mbar_init(&sc_full[s], GEMM_LN_PRODUCER_WARPS);
            mbar_init(&sc_empty[s], 1);
        }
        for (int s = 0; s < GEMM_LN_MAX_KB; ++s) {
            mbar_init(&a_full[s], 1);
            mbar_init(&a_empty[s], 1);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_pair<TMEM_COLS>(tmem_slot);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer: weight half tiles only
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int mt = pair; mt < m_tiles; mt += n_pairs) {
                for (int j = 0; j < n_tiles; ++j) {
                    const int n_blk = (j + n_rot) % n_tiles;
                    for (int kb = 0; kb < num_kb; ++kb) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * L::B_BYTES);
                        const uint32_t leader_full = mapa_shared(&full_bar[stage], 0);
                        tma_load_2d_pair(smem + L::B_OFFSET + stage * L::B_BYTES, &tmB, leader_full, kb * GEMM_BK,
                                         n_blk * BN + static_cast<int>(rank) * (BN / 2));
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (leader CTA only)
        if (rank == 0 && elect_one()) {
            constexpr uint32_t idesc = make_idesc_bf16(2 * GEMM_BM, BN, false);
            int stage = 0, as = 0;
            uint32_t phase = 0, aphase = 0, tphase = 0;
            long long pc_a = 0, pc_b = 0, pc_te = 0, pc_t = 0;
            const long long pc_start = p.prof ? clock64() : 0;
            for (int mt = pair; mt < m_tiles; mt += n_pairs, tphase ^= 1) {
                for (int j = 0; j < n_tiles; ++j) {
                    if (p.prof) pc_t = clock64();
                    mbar_wait(&tempty_bar[as], aphase ^ 1);
                    tc_fence_after();
                    if (p.prof) pc_te += clock64() - pc_t;
                    const uint32_t d_tmem = tmem_base + as * BN;
                    for (int kb = 0; kb < num_kb; ++kb) {
                        if (p.prof) pc_t = clock64();
                        if (j == 0) mbar_wait(&a_full[kb], tphase);          // both CTAs' k-block has landed (TMA bytes)
                        if (p.prof) { const long long t = clock64(); pc_a += t - pc_t; pc_t = t; }
                        mbar_wait(&full_bar[stage], phase);
                        tc_fence_after();
                        if (p.prof) pc_b += clock64() - pc_t;
                        const uint32_t a_addr = smem_u32(smem + kb * L::A_KB_BYTES);
                        const uint32_t b_addr = smem_u32(smem + L::B_OFFSET + stage * L::B_BYTES);
#pragma unroll
                        for (int k = 0; k < GEMM_BK / 16; ++k) {
                            const uint64_t ad = make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
                            const uint64_t bd = make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
                            umma_ss_pair(d_tmem, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
                        }
                        umma_commit_pair(&empty_bar[stage]);
                        if (j == n_tiles - 1) umma_commit_pair(&a_empty[kb]);  // last reader of this k-block: refill it
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                    umma_commit_pair(&tfull_bar[as]);
                    if (++as == 2) { as = 0; aphase ^= 1; }
                }
            }
            if (p.prof) {
                long long* o = p.prof + blockIdx.x * 16;
                o[0] = pc_a; o[1] = pc_b; o[2] = pc_te; o[3] = clock64() - pc_start;
            }
        }
    } else if (warp < GEMM_LN_W_PRODUCER) {
        // ------------------------------------------------------------------ epilogue (warps 2..9 of both CTAs)
        const int quarter = warp & 3;
        const int half = (warp - 2) >> 2;
        uint8_t* slab = smem + L::EPI_OFFSET + (warp - 2) * L::EPI_WARP_BYTES;
        int as = 0, buf = 0, tile_parity = 0, it = 0;
        uint32_t aphase = 0;
        const uint32_t leader_tempty0 = mapa_shared(&tempty_bar[0], 0), leader_tempty1 = mapa_shared(&tempty_bar[1], 0);
        for (int mt = pair; mt < m_tiles; mt += n_pairs, ++it) {
            const int m_blk = p.reverse ? m_tiles - 1 - mt : mt;
            for (int j = 0; j < n_tiles; ++j, ++tile_parity) {
                const int n_blk = (j + n_rot) % n_tiles;
                mbar_wait(&tfull_bar[as], aphase);
                tc_fence_after();
                // every k-block of this row tile has been multiplied at least once: its scratch buffer has been read in full
                if (j == 0 && warp == 2 && elect_one()) mbar_arrive(&sc_empty[it & 1]);
                const int m_warp = m_blk * 2 * GEMM_BM + static_cast<int>(rank) * GEMM_BM + quarter * 32;
                const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * BN;
                epilogue_store_tile<BN, 1>(p, tmC, tmCtail, slab, buf, t_row, m_warp, n_blk, half ^ (tile_parity & 1), lane);
                tc_fence_before();
                __syncwarp();
                if (elect_one()) mbar_arrive_cluster(as == 0 ? leader_tempty0 : leader_tempty1);
                if (++as == 2) { as = 0; aphase ^= 1; }
            }
        }
        bulk_wait<0>();
    } else if (warp < GEMM_LN_W_TMA_A) {
        // ------------------------------------------------------------------ LayerNorm producers (warps 10..17 of both CTAs)
        constexpr int ROWS_PER_WARP = GEMM_BM / GEMM_LN_PRODUCER_WARPS;      // 16
        constexpr int RB = GEMM_LN_ROWS_PER_BATCH;
        const int pw = warp - GEMM_LN_W_PRODUCER;
        const int D = p.K;
        const int vpl = D / 128;                                             // float4 vectors per lane (<= 3)
        int it = 0;
        long long pp_w = 0, pp_ld = 0, pp_t = 0, pp_r2 = 0, pp_st = 0, pp_f = 0;
        const long long pp_start = p.prof ? clock64() : 0;
        for (int mt = pair; mt < m_tiles; mt += n_pairs, ++it) {
            const int m_blk = p.reverse ? m_tiles - 1 - mt : mt;
            const int b = it & 1, u = it >> 1;
            if (p.prof) pp_t = clock64();
            if (u > 0) mbar_wait(&sc_empty[b], (u - 1) & 1);                 // TMA-A has read the tile that used this buffer
            if (p.prof) pp_w += clock64() - pp_t;
            const long long row0 = static_cast<long long>(m_blk) * 2 * GEMM_BM + static_cast<long long>(rank) * GEMM_BM + pw * ROWS_PER_WARP;
            __nv_bfloat16* dst0 = q.scratch + ((static_cast<long long>(blockIdx.x) * 2 + b) * GEMM_BM + pw * ROWS_PER_WARP) * D;
#pragma unroll 1
            for (int r0 = 0; r0 < ROWS_PER_WARP; r0 += RB) {
                if (q.debug & 1) break;
                // RB rows in flight per warp; every step below runs over all RB rows before the next one starts, so the
                // shuffle / divide / rsqrt latencies of the rows overlap (row after row, ptxas serialises the chains)
                float4 v[RB][3];
#pragma unroll
                for (int r = 0; r < RB; ++r) {
                    const long long row = row0 + r0 + r;
                    const float4* src = reinterpret_cast<const float4*>(q.x + row * q.ldx);
#pragma unroll
                    for (int i = 0; i < 3; ++i)
                        v[r][i] = (row < p.M && i < vpl) ? ((q.debug & 32) ? __ldg(src + lane + 32 * i) : ldg_stream_f4(src + lane + 32 * i)) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                float s[RB], mean[RB], rstd[RB];
                if (p.prof) pp_t = clock64();
#pragma unroll
                for (int r = 0; r < RB; ++r) {
                    s[r] = 0.f;
#pragma unroll
                    for (int i = 0; i < 3; ++i)
                        if (i < vpl) s[r] += (v[r][i].x + v[r][i].y) + (v[r][i].z + v[r][i].w);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                    for (int r = 0; r < RB; ++r) s[r] += __shfl_xor_sync(0xffffffffu, s[r], o);
                }
                if (p.prof) { const long long t = clock64(); pp_ld += t - pp_t; pp_t = t; }       // load latency + first reduction
#pragma unroll
                for (int r = 0; r < RB; ++r) {
                    mean[r] = s[r] / D;
                    float qq = 0.f;
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        if (i < vpl) {
                            const float a = v[r][i].x - mean[r], bb = v[r][i].y - mean[r], c = v[r][i].z - mean[r], d = v[r][i].w - mean[r];
                            qq += (a * a + bb * bb) + (c * c + d * d);
                        }
                    }
                    s[r] = qq;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                    for (int r = 0; r < RB; ++r) s[r] += __shfl_xor_sync(0xffffffffu, s[r], o);
                }
#pragma unroll
                for (int r = 0; r < RB; ++r) rstd[r] = rsqrtf(s[r] / D + q.eps);
                if (p.prof) { const long long t = clock64(); pp_r2 += t - pp_t; pp_t = t; }
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    if (i < vpl) {
                        const float4 g = __ldg(reinterpret_cast<const float4*>(q.gamma) + lane + 32 * i);
                        const float4 be = __ldg(reinterpret_cast<const float4*>(q.beta) + lane + 32 * i);
#pragma unroll
                        for (int r = 0; r < RB; ++r) {
                            float4 o;
                            o.x = (v[r][i].x - mean[r]) * rstd[r] * g.x + be.x;
                            o.y = (v[r][i].y - mean[r]) * rstd[r] * g.y + be.y;
                            o.z = (v[r][i].z - mean[r]) * rstd[r] * g.z + be.z;
                            o.w = (v[r][i].w - mean[r]) * rstd[r] * g.w + be.w;
                            uint2* dst = reinterpret_cast<uint2*>(dst0 + static_cast<long long>(r0 + r) * D);
                            dst[lane + 32 * i] = (row0 + r0 + r < p.M) ? make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w)) : make_uint2(0u, 0u);
                        }
                    }
                }
            }
            if (p.prof) { const long long t = clock64(); pp_st += t - pp_t; pp_t = t; }   // (last batch only: normalise + store)
            // the rows go from the generic proxy (st.global) to the async proxy (TMA-A's tensor load): make them visible at
            // GPU scope (the TMA unit reads L2), order them against the async proxy, then publish to the TMA-A thread
            if (!(q.debug & 2)) __threadfence();
            if (!(q.debug & 4)) fence_proxy_async_all();
            __syncwarp();
            if (elect_one()) mbar_arrive(&sc_full[b]);
            if (p.prof) pp_f += clock64() - pp_t;
        }
        if (p.prof && pw == 0 && lane == 0) {
            long long* o = p.prof + blockIdx.x * 16;
            o[4] = pp_w; o[5] = pp_ld; o[6] = clock64() - pp_start; o[10] = pp_r2; o[11] = pp_st; o[12] = pp_f;
        }
    } else {
        // ------------------------------------------------------------------ TMA-A: scratch tile -> stationary A k-blocks
        if (elect_one()) {
            int it = 0;
            long long pa_sc = 0, pa_e = 0, pa_t = 0;
            for (int mt = pair; mt < m_tiles; mt += n_pairs, ++it) {
                const int b = it & 1, u = it >> 1;
                if (p.prof) pa_t = clock64();
                mbar_wait(&sc_full[b], u & 1);
                if (p.prof) pa_sc += clock64() - pa_t;
                if (!(q.debug & 8)) fence_proxy_async_all();
                const int srow = (static_cast<int>(blockIdx.x) * 2 + b) * GEMM_BM;
                for (int kb = 0; kb < num_kb; ++kb) {
                    if (p.prof) pa_t = clock64();
                    if (it > 0) mbar_wait(&a_empty[kb], (it - 1) & 1);       // the previous row tile no longer reads this k-block
                    if (p.prof) pa_e += clock64() - pa_t;
                    if (rank == 0) mbar_expect_tx(&a_full[kb], 2 * L::A_KB_BYTES);
                    tma_load_2d_pair(smem + kb * L::A_KB_BYTES, &tmA, mapa_shared(&a_full[kb], 0), kb * GEMM_BK, srow);
                }
            }
            if (p.prof) { long long* o = p.prof + blockIdx.x * 16; o[8] = pa_sc; o[9] = pa_e; }
        }
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) tmem_dealloc_pair<TMEM_COLS>(tmem_base);
}

template <int BN>
static int launch_gemm2_ln(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmCtail,
                           const GemmParams& p, const GemmLnParams& q, int pairs, cudaStream_t stream) {
    using L = GemmLnSmem<BN>;
    B200X_TRY(ensure_kernel_smem(reinterpret_cast<const void*>(gemm2_ln_bf16_tn_kernel<BN>), L::TOTAL));
    gemm2_ln_bf16_tn_kernel<BN><<<2 * pairs, GEMM_LN_THREADS, L::TOTAL, stream>>>(tmA, tmB, tmC, tmCtail, p, q);
    B200X_CUDA_TRY(cudaGetLastError());
    return B200X_OK;
}

// bytes of the scratch area b200x_gemm_ln_bf16 needs on the current device: two normalised 128 x K bf16 row tiles per CTA
extern "C" int b200x_gemm_ln_scratch_bytes(int K, int64_t* bytes) {
    B200X_REQUIRE(bytes != nullptr && K > 0, "gemm_ln_scratch_bytes: bad argument");
    int num_sms = 0;
    B200X_TRY(device_sm_count(&num_sms));
    *bytes = static_cast<int64_t>(num_sms / 2) * 2 * 2 * GEMM_BM * K * 2;
    return B200X_OK;
}

extern "C" int b200x_gemm_ln_bf16(const float* d_x, int ldx, const float* d_gamma, const float* d_beta, float eps, const void* d_w,
                                  int ldw, int M, int N, int K, int block_n, void* d_out, int ldc, const float* d_bias, int act_gelu,
                                  void* d_scratch, int64_t scratch_bytes, int reverse, void* stream) {
    B200X_REQUIRE(d_x && d_gamma && d_beta && d_w && d_out && d_scratch, "gemm_ln: NULL argument");
    B200X_REQUIRE(M > 0 && N > 0 && N % 16 == 0, "gemm_ln: bad problem M=%d N=%d", M, N);
    B200X_REQUIRE(K % 128 == 0 && K <= GEMM_LN_MAX_KB * GEMM_BK, "gemm_ln: K=%d must be a multiple of 128 up to 384", K);
    B200X_REQUIRE(ldx % 4 == 0 && ldx >= K && (reinterpret_cast<uintptr_t>(d_x) & 15) == 0, "gemm_ln: x rows must be 16-byte aligned");
    B200X_REQUIRE((reinterpret_cast<uintptr_t>(d_gamma) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_beta) & 15) == 0, "gemm_ln: gamma / beta not 16-byte aligned");
    B200X_REQUIRE(d_bias == nullptr || (reinterpret_cast<uintptr_t>(d_bias) & 15) == 0, "gemm_ln: bias not 16-byte aligned");
    B200X_REQUIRE(ldw % 8 == 0 && ldc % 8 == 0, "gemm_ln: ldw=%d / ldc=%d must be multiples of 8", ldw, ldc);
    B200X_REQUIRE(block_n == 192 || block_n == 208 || block_n == 256, "gemm_ln: block_n %d unsupported (192/208/256)", block_n);
    B200X_REQUIRE((reinterpret_cast<uintptr_t>(d_scratch) & 1023) == 0, "gemm_ln: scratch must be 1024-byte aligned");
    int num_sms = 0;
    B200X_TRY(device_sm_count(&num_sms));
    const int pairs = std::min(ceil_div(M, 2 * GEMM_BM), num_sms / 2);
    const int64_t need = static_cast<int64_t>(pairs) * 2 * 2 * GEMM_BM * K * 2;
    B200X_REQUIRE(scratch_bytes >= need, "gemm_ln: scratch of %lld bytes, %lld needed (b200x_gemm_ln_scratch_bytes)",
                  static_cast<long long>(scratch_bytes), static_cast<long long>(need));
    CUtensorMap tmA, tmB, tmC, tmCtail;
    const uint64_t da[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(pairs) * 2 * 2 * GEMM_BM};
    const uint64_t sa[1] = {static_cast<uint64_t>(K) * 2};
    const uint32_t ba[2] = {GEMM_BK, GEMM_BM};
    B200X_TRY(make_tmap_bf16(&tmA, d_scratch, 2, da, sa, ba));
    const uint64_t dw[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(N)};
    const uint64_t sw[1] = {static_cast<uint64_t>(ldw) * 2};
    const uint32_t bw[2] = {GEMM_BK, static_cast<uint32_t>(block_n / 2)};
    B200X_TRY(make_tmap_bf16(&tmB, d_w, 2, dw, sw, bw));
    const uint64_t dc[2] = {static_cast<uint64_t>(N), static_cast<uint64_t>(M)};
    const uint64_t sc[1] = {static_cast<uint64_t>(ldc) * 2};
    const uint32_t bc[2] = {64, 32}, bt[2] = {16, 32};
    B200X_TRY(make_tmap(&tmC, d_out, 2, 2, dc, sc, bc, 1));
    B200X_TRY(make_tmap(&tmCtail, d_out, 2, 2, dc, sc, bt, 0));
    GemmParams p{M, N, K, d_out, ldc, B200X_GEMM_OUT_BF16, d_bias, act_gelu, nullptr, 0, 0, 0, nullptr, reverse ? 1 : 0};
    GemmLnParams q{d_x, ldx, d_gamma, d_beta, eps, reinterpret_cast<__nv_bfloat16*>(d_scratch), getenv("B200X_GEMM_LN_DEBUG") ? atoi(getenv("B200X_GEMM_LN_DEBUG")) : 0};
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    static int prof_left = getenv("B200X_GEMM_LN_PROF") ? atoi(getenv("B200X_GEMM_LN_PROF")) : 0;
    long long* d_prof = nullptr;
    if (prof_left > 0) {
        cudaMalloc(&d_prof, sizeof(long long) * 16 * 2 * pairs);
        cudaMemset(d_prof, 0, sizeof(long long) * 16 * 2 * pairs);
        p.prof = d_prof;
    }
    int rc;
    switch (block_n) {
        case 192: rc = launch_gemm2_ln<192>(tmA, tmB, tmC, tmCtail, p, q, pairs, s); break;
        case 208: rc = launch_gemm2_ln<208>(tmA, tmB, tmC, tmCtail, p, q, pairs, s); break;
        default: rc = launch_gemm2_ln<256>(tmA, tmB, tmC, tmCtail, p, q, pairs, s); break;
    }
    if (d_prof != nullptr) {
        --prof_left;
        cudaDeviceSynchronize();
        std::vector<long long> h(16 * 2 * pairs);
        cudaMemcpy(h.data(), d_prof, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
        cudaFree(d_prof);
        double a[16] = {0};
        for (int c = 0; c < 2 * pairs; ++c) for (int i = 0; i < 16; ++i) a[i] += static_cast<double>(h[c * 16 + i]);
        const double tiles = static_cast<double>(ceil_div(M, 2 * GEMM_BM)) / pairs;
        fprintf(stderr, "gemm_ln prof N=%d: row tiles/pair %.1f | issuer (leader, per CTA-pair): wait a_full %.0f  wait B %.0f  wait tempty %.0f  total %.0f clk"
                " | producer warp 0 (per CTA): wait sc_empty %.0f  load+reduce1 %.0f  var+reduce2 %.0f  last-batch store %.0f  fence %.0f  total %.0f | TMA-A: wait sc_full %.0f  wait a_empty %.0f\n",
                N, tiles, a[0] / pairs, a[1] / pairs, a[2] / pairs, a[3] / pairs, a[4] / (2 * pairs), a[5] / (2 * pairs), a[10] / (2 * pairs), a[11] / (2 * pairs), a[12] / (2 * pairs), a[6] / (2 * pairs),
                a[8] / (2 * pairs), a[9] / (2 * pairs));
    }
    return rc;
}
