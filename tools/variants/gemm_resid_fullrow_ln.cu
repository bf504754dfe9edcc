// REJECTED VARIANT (measured, not part of libb200xai.so): residual GEMM + LayerNorm with the WHOLE 384-column row in TMEM.
// x is read once by TMA, x_new goes back to HBM and into its TMEM columns, row statistics per thread (TMEM lane == row), h by a
// second TMEM pass.  Correct (x_new bit-identical to the reduce-add path, h within one bf16 ulp of the LayerNorm pass on < 2 % of
// the elements; tests passed), but with both accumulator tiles holding one row tile the MMAs wait for the epilogue, the
// epilogue's per-warp chain (TMA in -> TMEM -> slab -> TMA out, six groups, then two more TMEM passes) takes ~22 k cycles per
// row tile, and its 16 KB of slabs per warp leave room for only three ring stages (1 050 cycles per k-step instead of ~450):
//   229 copies (profiles/r02_j_gemm_resid_fullrow.txt):  proj 368 us (residual GEMM 210 + LayerNorm 107 = 317 us separately)
//                                                         fc2  616 us (309 + 107 = 416 us separately; 370-390 us with the tail)
// This file is an excerpt (it needs the helpers of csrc/gemm_tcgen05.cu around it to compile).

// ------------------------------------------------------------------------------------------------ full-row residual GEMM + LayerNorm
// x_new = x + A . W^T + bias and h = LayerNorm(x_new) in ONE pass over x: the accumulators of BOTH column tiles of a row tile
// (N = 384 = 2 x 192 fp32 columns) stay in TMEM, so an epilogue thread (TMEM lane == row) sees its whole row:
//   pass 1 : x_old arrives by TMA (32 x 32 fp32 boxes, prefetched into L2 one row tile ahead), v = (acc + bias) + x_old - the
//            same sum the TMA reduce-add of the plain residual epilogue forms, bit for bit - goes back to HBM by TMA store
//            from the slab it arrived in, and back into its TMEM columns (tcgen05.st); row sums accumulate in registers
//   pass 1b: mean -> sum of squared deviations from TMEM (two-pass variance, no cancellation)
//   pass 2 : (v - mean) * rstd * gamma + beta -> bf16 -> slab -> TMA store of h
// Two warps share a TMEM lane quarter and split the row's six 64-column groups; their partial sums meet in shared memory
// behind a 64-thread named barrier.  No accumulator double buffering: the MMAs of the next row tile wait for pass 2 - these
// GEMMs are HBM-bound (x read + x write + h write + A), the tensor pipe has the slack.
// HBM bytes per call at 229 copies: 1.45 GB (proj) / 1.87 GB (fc2) against 1.94 / 2.35 GB for residual GEMM + LayerNorm pass.
constexpr int GEMM_RFL_BN = 192;
constexpr int GEMM_RFL_XSLABS = 4;                        // 4 KB slabs per epilogue warp (x_old in / x_new out / h out)

struct GemmRflSmem {
    static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
    static constexpr int B_BYTES = (GEMM_RFL_BN / 2) * GEMM_BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int EPI_WARP_BYTES = GEMM_RFL_XSLABS * GEMM_SLAB_BYTES;
    static constexpr int STAT_BYTES = 2 * 2 * GEMM_EPI_WARPS * 32 * 4;       // [row tile parity][sum | sum of squares][warp][lane]
    static constexpr int STAGES = (232448 - 1024 - 768 - STAT_BYTES - GEMM_EPI_WARPS * EPI_WARP_BYTES) / STAGE_BYTES;
    static constexpr int EPI_OFFSET = STAGES * STAGE_BYTES;
    static constexpr int STAT_OFFSET = EPI_OFFSET + GEMM_EPI_WARPS * EPI_WARP_BYTES;
    static constexpr int BAR_OFFSET = STAT_OFFSET + STAT_BYTES;
    static constexpr int TOTAL = BAR_OFFSET + 768 + 1024;
    static_assert(STAGES >= 3 && TOTAL <= 232448, "shared memory budget exceeded");
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm2_resid_fullrow_ln_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                              const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmH, GemmParams p,
                              GemmLnTail q) {
    using L = GemmRflSmem;
    constexpr int BN = GEMM_RFL_BN;
    constexpr int STAGES = L::STAGES;
    constexpr int XS = GEMM_RFL_XSLABS;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;            // [2] column tile n of the row tile is in TMEM
    uint64_t* tempty_bar = tfull_bar + 2;                // [1] the row tile's accumulators have been drained (leader's copy)
    uint64_t* xfull_bar = tempty_bar + 1;                // [8 warps][XS] x_old box has landed in the warp's slab
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xfull_bar + GEMM_EPI_WARPS * XS);
    float* stat = reinterpret_cast<float*>(smem + L::STAT_OFFSET);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int m_tiles = (p.M + 2 * GEMM_BM - 1) / (2 * GEMM_BM);
    const int num_kb = (p.K + GEMM_BK - 1) / GEMM_BK;

    if (warp == 0 && elect_one()) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmH);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(&tfull_bar[0], 1);
        mbar_init(&tfull_bar[1], 1);
        mbar_init(tempty_bar, 2 * GEMM_EPI_WARPS);
        for (int s = 0; s < GEMM_EPI_WARPS * XS; ++s) mbar_init(&xfull_bar[s], 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_pair<512>(tmem_slot);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer (both CTAs)
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int mt = pair; mt < m_tiles; mt += n_pairs) {
                const int m_blk = p.reverse ? m_tiles - 1 - mt : mt;
                for (int n_blk = 0; n_blk < 2; ++n_blk) {
                    for (int kb = 0; kb < num_kb; ++kb) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        uint8_t* sa = smem + stage * L::STAGE_BYTES;
                        if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * L::STAGE_BYTES);
                        const uint32_t leader_full = mapa_shared(&full_bar[stage], 0);
                        tma_load_2d_pair(sa, &tmA, leader_full, kb * GEMM_BK, m_blk * 2 * GEMM_BM + static_cast<int>(rank) * GEMM_BM);
                        tma_load_2d_pair(sa + L::A_BYTES, &tmB, leader_full, kb * GEMM_BK, n_blk * BN + static_cast<int>(rank) * (BN / 2));
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (leader CTA only)
        if (rank == 0 && elect_one()) {
            constexpr uint32_t idesc = make_idesc_bf16(2 * GEMM_BM, BN, false);
            int stage = 0;
            uint32_t phase = 0, tphase = 0;
            for (int mt = pair; mt < m_tiles; mt += n_pairs, tphase ^= 1) {
                mbar_wait(tempty_bar, tphase ^ 1);       // both CTAs' epilogues are done with the previous row tile's TMEM
                tc_fence_after();
                for (int n_blk = 0; n_blk < 2; ++n_blk) {
                    const uint32_t d_tmem = tmem_base + n_blk * BN;
                    for (int kb = 0; kb < num_kb; ++kb) {
                        mbar_wait(&full_bar[stage], phase);
                        tc_fence_after();
                        const uint32_t a_addr = smem_u32(smem + stage * L::STAGE_BYTES);
                        const uint32_t b_addr = a_addr + L::A_BYTES;
#pragma unroll
                        for (int k = 0; k < GEMM_BK / 16; ++k) {
                            const uint64_t ad = make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
                            const uint64_t bd = make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
                            umma_ss_pair(d_tmem, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
                        }
                        umma_commit_pair(&empty_bar[stage]);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                    umma_commit_pair(&tfull_bar[n_blk]);
                }
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (warps 2..9 of both CTAs)
        const int we = warp - 2;
        const int quarter = warp & 3;                    // TMEM lanes [32 quarter, 32 quarter + 32) = rows of the CTA's tile
        const int hf = we >> 2;                          // which of the two warps of the quarter: 64-column groups S = 2 i + hf
        const int partner = (we & 3) | ((hf ^ 1) << 2);  // the other warp of this quarter (epilogue warp index)
        const uint32_t bar_id = 1 + (we & 3);            // named barrier of the quarter's two warps (64 threads)
        uint8_t* slab = smem + L::EPI_OFFSET + we * L::EPI_WARP_BYTES;
        uint64_t* xfull = xfull_bar + we * XS;
        const uint32_t leader_tempty = mapa_shared(tempty_bar, 0);
        const int sw = lane & 7;
        const float inv_d = 1.0f / static_cast<float>(p.N);
        uint32_t xph = 0;                                // bit s: parity of the next completion of xfull[s]
        uint32_t tphase = 0;
        int it = 0;
        // column of group j (0..5) of this warp: 64-column group S = 2 (j >> 1) + hf, 32-column half j & 1
        auto col_of = [&](int j) { return 64 * (2 * (j >> 1) + hf) + 32 * (j & 1); };
        auto row_of = [&](int mt) {
            const int m_blk = p.reverse ? m_tiles - 1 - mt : mt;
            return m_blk * 2 * GEMM_BM + static_cast<int>(rank) * GEMM_BM + quarter * 32;
        };
        auto load_x = [&](int j, int m_warp) {           // elected lane: x_old box of group j -> slab j % XS
            mbar_expect_tx(&xfull[j % XS], GEMM_SLAB_BYTES);
            tma_load_2d(slab + (j % XS) * GEMM_SLAB_BYTES, &tmX, &xfull[j % XS], col_of(j), m_warp);
        };
        if (pair < m_tiles && elect_one()) {
            const int m_warp = row_of(pair);
#pragma unroll
            for (int j = 0; j < XS; ++j) load_x(j, m_warp);
        }
        for (int mt = pair; mt < m_tiles; mt += n_pairs, ++it, tphase ^= 1) {
            const int m_warp = row_of(mt);
            const bool more = mt + n_pairs < m_tiles;
            if (more && elect_one()) {                   // the next row tile's x_old on its way into L2 (a tile time ahead)
                const int m_next = row_of(mt + n_pairs);
#pragma unroll
                for (int j = 0; j < 6; ++j) tma_prefetch_l2_2d(&tmX, col_of(j), m_next);
            }
            const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
            // ---- pass 1: v = (acc + bias) + x_old -> HBM, TMEM; row sum
            float s1a = 0.f, s1b = 0.f;
#pragma unroll 1
            for (int j = 0; j < 6; ++j) {
                const int c = col_of(j);
                if (j == 0) { mbar_wait(&tfull_bar[0], tphase); tc_fence_after(); }
                if (c >= BN && col_of(j - 1) < BN) { mbar_wait(&tfull_bar[1], tphase); tc_fence_after(); }
                const int sl = j % XS;
                mbar_wait(&xfull[sl], (xph >> sl) & 1);
                xph ^= 1u << sl;
                uint8_t* xs = slab + sl * GEMM_SLAB_BYTES;
                uint32_t r[32];
                tmem_ld32(t_row + c, r);
                tmem_wait_ld();
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (p.bias != nullptr) b = __ldg(reinterpret_cast<const float4*>(p.bias + c + 4 * k));
                    float4* cell = reinterpret_cast<float4*>(xs + lane * 128 + ((k ^ sw) << 4));
                    const float4 xo = *cell;
                    float4 v;
                    v.x = (__uint_as_float(r[4 * k]) + b.x) + xo.x;
                    v.y = (__uint_as_float(r[4 * k + 1]) + b.y) + xo.y;
                    v.z = (__uint_as_float(r[4 * k + 2]) + b.z) + xo.z;
                    v.w = (__uint_as_float(r[4 * k + 3]) + b.w) + xo.w;
                    *cell = v;
                    r[4 * k] = __float_as_uint(v.x); r[4 * k + 1] = __float_as_uint(v.y);
                    r[4 * k + 2] = __float_as_uint(v.z); r[4 * k + 3] = __float_as_uint(v.w);
                    s1a += v.x + v.y;
                    s1b += v.z + v.w;
                }
                tmem_st32(t_row + c, r);
                fence_proxy_async();
                __syncwarp();
                if (elect_one()) {
                    tma_store_2d(&tmX, xs, c, m_warp);   // rows >= M are clipped by TMA
                    bulk_commit();
                    if (j + XS - 1 < 6 && j >= 1) {      // groups 4, 5 reuse slabs 0, 1: their stores (groups 0, 1) have been read
                        bulk_wait_read<1>();
                        load_x(j + XS - 1, m_warp);
                    }
                }
            }
            tmem_wait_st();
            // ---- row mean: the quarter's two warps exchange their partial sums
            float* st_sum = stat + ((it & 1) * 2 + 0) * GEMM_EPI_WARPS * 32;
            float* st_sq = stat + ((it & 1) * 2 + 1) * GEMM_EPI_WARPS * 32;
            const float my_sum = s1a + s1b;
            st_sum[we * 32 + lane] = my_sum;
            named_bar_sync(bar_id, 64);
            const float other_sum = st_sum[partner * 32 + lane];
            const float mean = ((hf == 0) ? (my_sum + other_sum) : (other_sum + my_sum)) * inv_d;
            // ---- pass 1b: sum of squared deviations, from TMEM
            float s2a = 0.f, s2b = 0.f;
#pragma unroll 1
            for (int j = 0; j < 6; ++j) {
                uint32_t r[32];
                tmem_ld32(t_row + col_of(j), r);
                tmem_wait_ld();
#pragma unroll
                for (int k = 0; k < 32; k += 2) {
                    const float d0 = __uint_as_float(r[k]) - mean, d1 = __uint_as_float(r[k + 1]) - mean;
                    s2a = fmaf(d0, d0, s2a);
                    s2b = fmaf(d1, d1, s2b);
                }
            }
            const float my_sq = s2a + s2b;
            st_sq[we * 32 + lane] = my_sq;
            named_bar_sync(bar_id, 64);
            const float other_sq = st_sq[partner * 32 + lane];
            const float rstd = rsqrtf(fmaf((hf == 0) ? (my_sq + other_sq) : (other_sq + my_sq), inv_d, q.eps));
            // ---- pass 2: normalise -> bf16 -> slab (128-byte rows of 64 columns) -> TMA store of h
#pragma unroll 1
            for (int g = 0; g < 3; ++g) {
                const int c = 64 * (2 * g + hf);
                uint8_t* hs = slab + (g & 1 ? 3 : 2) * GEMM_SLAB_BYTES;      // slabs 2 / 3 (no x_old load is ever pending on them here)
                uint32_t r[64];
                tmem_ld32(t_row + c, r);
                tmem_ld32(t_row + c + 32, r + 32);
                tmem_wait_ld();
                bulk_wait_read<1>();                     // the store that last used this slab has read it
                __syncwarp();
#pragma unroll
                for (int k = 0; k < 8; ++k) {            // 8 columns -> one 16-byte chunk
                    const float4 g0 = __ldg(reinterpret_cast<const float4*>(q.gamma + c + 8 * k));
                    const float4 g1 = __ldg(reinterpret_cast<const float4*>(q.gamma + c + 8 * k + 4));
                    const float4 b0 = __ldg(reinterpret_cast<const float4*>(q.beta + c + 8 * k));
                    const float4 b1 = __ldg(reinterpret_cast<const float4*>(q.beta + c + 8 * k + 4));
                    const float o0 = (__uint_as_float(r[8 * k]) - mean) * rstd * g0.x + b0.x;
                    const float o1 = (__uint_as_float(r[8 * k + 1]) - mean) * rstd * g0.y + b0.y;
                    const float o2 = (__uint_as_float(r[8 * k + 2]) - mean) * rstd * g0.z + b0.z;
                    const float o3 = (__uint_as_float(r[8 * k + 3]) - mean) * rstd * g0.w + b0.w;
                    const float o4 = (__uint_as_float(r[8 * k + 4]) - mean) * rstd * g1.x + b1.x;
                    const float o5 = (__uint_as_float(r[8 * k + 5]) - mean) * rstd * g1.y + b1.y;
                    const float o6 = (__uint_as_float(r[8 * k + 6]) - mean) * rstd * g1.z + b1.z;
                    const float o7 = (__uint_as_float(r[8 * k + 7]) - mean) * rstd * g1.w + b1.w;
                    *reinterpret_cast<uint4*>(hs + lane * 128 + ((k ^ sw) << 4)) =
                        make_uint4(pack_bf16(o0, o1), pack_bf16(o2, o3), pack_bf16(o4, o5), pack_bf16(o6, o7));
                }
                fence_proxy_async();
                __syncwarp();
                if (elect_one()) {
                    tma_store_2d(&tmH, hs, c, m_warp);
                    bulk_commit();
                }
            }
            // ---- TMEM is free for the next row tile; fetch its first x_old boxes behind the stores that used the slabs
            tc_fence_before();
            __syncwarp();
            if (elect_one()) {
                mbar_arrive_cluster(leader_tempty);
                if (more) {
                    bulk_wait_read<0>();
                    const int m_next = row_of(mt + n_pairs);
#pragma unroll
                    for (int j = 0; j < XS; ++j) load_x(j, m_next);
                }
            }
        }
        bulk_wait<0>();
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) tmem_dealloc_pair<512>(tmem_base);
}

/* Full-row variant (gemm2_resid_fullrow_ln_kernel): N must be exactly 384. */
extern "C" int b200x_gemm_resid_ln_fullrow_bf16(const void* d_a, int lda, const void* d_w, int ldw, int M, int N, int K, float* d_x,
                                                int ldx, const float* d_bias, const float* d_gamma, const float* d_beta, float eps,
                                                void* d_h, int ldh, int reverse, void* stream) {
    B200X_REQUIRE(d_a && d_w && d_x && d_gamma && d_beta && d_h, "gemm_resid_ln_fullrow: NULL argument");
    B200X_REQUIRE(M > 0 && K > 0, "gemm_resid_ln_fullrow: empty problem");
    B200X_REQUIRE(N == 2 * GEMM_RFL_BN, "gemm_resid_ln_fullrow: N=%d, the kernel holds one 384-column row per TMEM lane", N);
    B200X_REQUIRE(K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0, "gemm_resid_ln_fullrow: K / lda / ldw must be multiples of 8");
    B200X_REQUIRE(ldx % 4 == 0 && ldx >= N && ldh % 8 == 0 && ldh >= N, "gemm_resid_ln_fullrow: ldx / ldh misaligned");
    B200X_REQUIRE((reinterpret_cast<uintptr_t>(d_x) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_h) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(d_gamma) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_beta) & 15) == 0 &&
                  (d_bias == nullptr || (reinterpret_cast<uintptr_t>(d_bias) & 15) == 0), "gemm_resid_ln_fullrow: misaligned pointer");
    CUtensorMap tmA, tmB, tmX, tmH;
    const uint64_t da[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(M)};
    const uint64_t sa[1] = {static_cast<uint64_t>(lda) * 2};
    const uint32_t ba[2] = {GEMM_BK, GEMM_BM};
    B200X_TRY(make_tmap_bf16(&tmA, d_a, 2, da, sa, ba));
    const uint64_t dw[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(N)};
    const uint64_t sw[1] = {static_cast<uint64_t>(ldw) * 2};
    const uint32_t bw[2] = {GEMM_BK, GEMM_RFL_BN / 2};
    B200X_TRY(make_tmap_bf16(&tmB, d_w, 2, dw, sw, bw));
    const uint64_t dx[2] = {static_cast<uint64_t>(N), static_cast<uint64_t>(M)};
    const uint64_t sx[1] = {static_cast<uint64_t>(ldx) * 4};
    const uint32_t bx[2] = {32, 32};
    B200X_TRY(make_tmap(&tmX, d_x, 4, 2, dx, sx, bx, 1));
    const uint64_t sh[1] = {static_cast<uint64_t>(ldh) * 2};
    const uint32_t bh[2] = {64, 32};
    B200X_TRY(make_tmap(&tmH, d_h, 2, 2, dx, sh, bh, 1));
    GemmParams p{M, N, K, d_x, ldx, B200X_GEMM_OUT_F32_RESID, d_bias, 0, nullptr, 0, 0, 0, nullptr, reverse ? 1 : 0};
    GemmLnTail q{d_x, ldx, d_gamma, d_beta, eps, reinterpret_cast<__nv_bfloat16*>(d_h), ldh};
    int num_sms = 0;
    B200X_TRY(ensure_kernel_smem(reinterpret_cast<const void*>(gemm2_resid_fullrow_ln_kernel), GemmRflSmem::TOTAL));
    B200X_TRY(device_sm_count(&num_sms));
    const int pairs = std::min(ceil_div(M, 2 * GEMM_BM), num_sms / 2);
    gemm2_resid_fullrow_ln_kernel<<<2 * pairs, GEMM_THREADS, GemmRflSmem::TOTAL, static_cast<cudaStream_t>(stream)>>>(tmA, tmB, tmX, tmH, p, q);
    B200X_CUDA_TRY(cudaGetLastError());
    return B200X_OK;
}
