// REJECTED VARIANT (measured, not part of libb200xai.so): 64-key tiles with THREE CTAs per SM and P in its own TMEM columns
// (S 64 + O 64 in a 128-column allocation, P in a second 32-column allocation: 160 columns per CTA, 3 x 64 KB of shared memory).
// The round-1 64-key kernel aliased P into S (strictly serial chains); this one keeps S(j+1) overlapped with the exponentials
// of tile j like the production kernel and adds a third independent chain per SM.  Correct (13 attention tests pass), but
// 1 031 us against 992 us at 229 copies (300.9 vs 288.8 us at 64): twice as many barrier round trips per key, N = 64 UMMAs that
// re-read the Q operand from shared memory twice as often, and 96 registers per thread (56 bytes spilled) cost more than the
// third chain gains (profiles/r02_q_attention_k64x3.txt).
// This file is an excerpt of csrc/attention_tcgen05.cu (compile there with -DB200X_ATT_K64; needs tmem_alloc_keep in ptx.cuh).

#ifdef B200X_ATT_K64
// ------------------------------------------------------------------------------------------------ experiment: 64-key tiles, THREE CTAs per SM
// S (64 columns) + O (64) in one 128-column TMEM allocation, P (32 columns) in a second one: 160 columns per CTA, three CTAs per SM
// (480 of 512 columns, 3 x 64 KB of shared memory, 112 registers): a third independent S -> softmax -> PV chain per SM.
constexpr int ATT64_KEYS = 64;
constexpr int ATT64_KV_BYTES = ATT64_KEYS * ATT_HD * 2;       // 8 KB
constexpr int ATT64_STAGES = 3;
constexpr int ATT64_SMEM = ATT_TILE_BYTES + 2 * ATT64_STAGES * ATT64_KV_BYTES + 256;

__global__ void __launch_bounds__(ATT_THREADS, 3)
attention_k64_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, AttnParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sQ = smem;
    uint8_t* sK = smem + ATT_TILE_BYTES;
    uint8_t* sV = sK + ATT64_STAGES * ATT64_KV_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + ATT64_STAGES * ATT64_KV_BYTES);
    uint64_t* q_full = bars;
    uint64_t* kv_full = bars + 1;
    uint64_t* kv_empty = kv_full + ATT64_STAGES;
    uint64_t* s_full = kv_empty + ATT64_STAGES;
    uint64_t* s_free = s_full + 1;
    uint64_t* p_ready = s_free + 1;
    uint64_t* pv_done = p_ready + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 1);      // [0]: S | O (128 columns), [1]: P (32 columns)

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int head = p.reverse ? gridDim.y - 1 - blockIdx.y : blockIdx.y;
    const int copy = p.reverse ? gridDim.z - 1 - blockIdx.z : blockIdx.z;
    const int q0 = blockIdx.x * ATT_TILE;
    const int nkv = (p.tokens + ATT64_KEYS - 1) / ATT64_KEYS;
    const int hidden = p.heads * ATT_HD;

    if (warp == ATT_W_TMA && elect_one()) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmKV);
        mbar_init(q_full, 1);
        for (int s = 0; s < ATT64_STAGES; ++s) {
            mbar_init(&kv_full[s], 1);
            mbar_init(&kv_empty[s], 1);
        }
        mbar_init(s_full, 1);
        mbar_init(s_free, 4);
        mbar_init(p_ready, 4);
        mbar_init(pv_done, 1);
        fence_barrier_init();
    }
    if (warp == ATT_W_MMA) {
        tmem_alloc_keep<128>(tmem_slot);
        tmem_alloc<32>(tmem_slot + 1);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_so = tmem_slot[0], tmem_p = tmem_slot[1];

    if (warp == ATT_W_TMA) {
        if (elect_one()) {
            mbar_expect_tx(q_full, ATT_TILE_BYTES);
            tma_load_3d(sQ, &tmQ, q_full, head * ATT_HD, q0, copy);
            for (int j = 0; j < nkv; ++j) {
                const int st = j % ATT64_STAGES;
                mbar_wait(&kv_empty[st], ((j / ATT64_STAGES) & 1) ^ 1);
                mbar_expect_tx(&kv_full[st], 2 * ATT64_KV_BYTES);
                tma_load_3d(sK + st * ATT64_KV_BYTES, &tmKV, &kv_full[st], hidden + head * ATT_HD, j * ATT64_KEYS, copy);
                tma_load_3d(sV + st * ATT64_KV_BYTES, &tmKV, &kv_full[st], 2 * hidden + head * ATT_HD, j * ATT64_KEYS, copy);
            }
        }
    } else if (warp == ATT_W_MMA) {
        if (elect_one()) {
            constexpr uint32_t idesc_pv = make_idesc_bf16(ATT_TILE, ATT_HD, true);
            const uint32_t tS = tmem_so, tO = tmem_so + 64, tP = tmem_p;
            const uint64_t q_desc = make_smem_desc_sw128(smem_u32(sQ), 16, 1024);
            const uint64_t k_desc0 = make_smem_desc_sw128(smem_u32(sK), 16, 1024);
            const uint64_t v_desc0 = make_smem_desc_sw128(smem_u32(sV), 8192, 1024);
            auto issue_s = [&](int j) {
                const int nk = min(ATT64_KEYS, p.tokens - j * ATT64_KEYS);
                const uint32_t idesc_s = make_idesc_bf16(ATT_TILE, nk, false);
                const uint64_t kd = k_desc0 + static_cast<uint64_t>((j % ATT64_STAGES) * (ATT64_KV_BYTES >> 4));
#pragma unroll
                for (int k = 0; k < ATT_HD / 16; ++k) umma_ss(tS, q_desc + 2 * k, kd + 2 * k, idesc_s, k != 0 ? 1u : 0u);
                umma_commit(s_full);
            };
            auto issue_pv = [&](int j) {
                const int nk = min(ATT64_KEYS, p.tokens - j * ATT64_KEYS);
                const uint64_t vd = v_desc0 + static_cast<uint64_t>((j % ATT64_STAGES) * (ATT64_KV_BYTES >> 4));
                for (int ks = 0; ks < nk / 16; ++ks) umma_ts(tO, tP + ks * 8, vd + 128 * ks, idesc_pv, (j | ks) != 0 ? 1u : 0u);
                umma_commit(pv_done);
                umma_commit(&kv_empty[j % ATT64_STAGES]);
            };
            mbar_wait(q_full, 0);
            mbar_wait(&kv_full[0], 0);
            tc_fence_after();
            issue_s(0);
            for (int j = 0; j < nkv; ++j) {
                if (j + 1 < nkv) {
                    mbar_wait(&kv_full[(j + 1) % ATT64_STAGES], ((j + 1) / ATT64_STAGES) & 1);
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
        const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
        const uint32_t tS = tmem_so + lane_off, tO = tmem_so + lane_off + 64, tP = tmem_p + lane_off;
        const float c = p.scale_log2;
        const uint64_t c2 = pack_f32x2(c, c);
        const uint64_t zero2 = pack_f32x2(p.zero, p.zero);
        float m_ref = -INFINITY;
        uint64_t l2 = 0ull, l2b = 0ull;
        uint32_t r[ATT64_KEYS];
        for (int j = 0; j < nkv; ++j) {
            const int nk = min(ATT64_KEYS, p.tokens - j * ATT64_KEYS);
            mbar_wait(s_full, j & 1);
            tc_fence_after();
            if (nk == ATT64_KEYS) {
                tmem_ld32(tS, r); tmem_ld32(tS + 32, r + 32);
            } else {
#pragma unroll
                for (int col = 0; col < ATT64_KEYS; col += 16) {
                    if (col < nk) {
                        tmem_ld16(tS + col, r + col);
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i) r[col + i] = 0xff800000u;
                    }
                }
            }
            tmem_wait_ld();
            tc_fence_before();
            __syncwarp();
            if (elect_one()) mbar_arrive(s_free);
            float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
            for (int i = 0; i < ATT64_KEYS; i += 8) {
                m0 = fmax3(m0, __uint_as_float(r[i]), __uint_as_float(r[i + 1]));
                m1 = fmax3(m1, __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
                m2 = fmax3(m2, __uint_as_float(r[i + 4]), __uint_as_float(r[i + 5]));
                m3 = fmax3(m3, __uint_as_float(r[i + 6]), __uint_as_float(r[i + 7]));
            }
            const float mt = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
            if (j == 0) m_ref = mt;
            const bool need = (mt - m_ref) * c > ATT_RESCALE_LOG2;
            bool pv_waited = false;
            if (__any_sync(0xffffffffu, need)) {
                if (j > 0) { mbar_wait(pv_done, (j - 1) & 1); tc_fence_after(); pv_waited = true; }
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
            if (j > 0 && !pv_waited) { mbar_wait(pv_done, (j - 1) & 1); tc_fence_after(); }
            tmem_st16(tP, pk);
            exp_chunk(r + 32, pk, c2, nmc2, zero2, l2, l2b);
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
    if (warp == ATT_W_MMA) {
        tmem_dealloc<128>(tmem_so);
        tmem_dealloc<32>(tmem_p);
    }
}
#endif


// ---- launcher fragment (inside b200x_attention) ----
#ifdef B200X_ATT_K64
    {
        CUtensorMap tmKV;
        const uint32_t boxkv[3] = {ATT_HD, ATT64_KEYS, 1};
        B200X_TRY(make_tmap_bf16(&tmKV, d_qkv, 3, dims, strides, boxkv));
        B200X_TRY(ensure_kernel_smem(reinterpret_cast<const void*>(attention_k64_kernel), ATT64_SMEM, true));
        dim3 grid64(ceil_div(tokens, ATT_TILE), heads, copies);
        attention_k64_kernel<<<grid64, ATT_THREADS, ATT64_SMEM, static_cast<cudaStream_t>(stream)>>>(tm, tmKV, p);
        B200X_CUDA_TRY(cudaGetLastError());
        return B200X_OK;
    }
#endif
