// Persistent, warp-specialised bf16 GEMM for sm_100a:  C[M,N] = A[M,K] . W[N,K]^T  (+ fused epilogue)
//
//   warp 0    : TMA producer (cp.async.bulk.tensor, 128B swizzle, 3-stage mbarrier ring)
//   warp 1    : TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128 x BN x 16, fp32 accumulate in TMEM)
//   warps 2-9 : epilogue.  TMEM lane == tile row, so tcgen05.ld hands every thread one row of the accumulator.
//               bf16 / residual outputs: bias (+GELU) in registers, pack, st.shared into a 128B-swizzled per-warp slab
//               and ONE TMA tensor store per 32-row x 128-byte slab - a plain store for bf16 outputs, a TMA reduce-add
//               (x += acc + bias performed in L2) for the fp32 residual stream, which therefore is never loaded.
//               token-mode outputs (tokenizer: row remap + positional encoding) take a generic coalescing path.
// Two TMEM accumulator stages let the epilogue of tile i overlap the MMAs of tile i+1.  With K = 384 the epilogue is as
// long as the MMAs, hence eight epilogue warps (two per TMEM lane quarter, splitting the column groups).
// What bounds it (in-kernel cycle counters, tools/gemm_profile.py): a 128 x BN tile pulls (16 KB + BN x 128 B) per 64-deep
// k-block through the SM's L2 port, which delivers ~54 B/clk; for BN = 192 that is 40 KB per 424 MMA cycles, so the port -
// not the tensor pipe, not the ring depth (3 vs 4 stages measured equal) - paces the kernel at ~50 % of the MMA rate.
// Hence the CTA-pair kernel below (cta_group::2: half of the B tile per SM) for everything but the tokenizer epilogue.
// Used for every dense contraction of the SpecTTTra forward (tokenizer projections, QKV, attention projection, MLP).
#include <algorithm>
#include <type_traits>

#include "common.h"
#include "ptx.cuh"
#include "layernorm_rows.cuh"

namespace b200x {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;      // 64 bf16 = one 128-byte swizzle row
constexpr int GEMM_MAX_STAGES = 3;
constexpr int GEMM_EPI_WARPS = 8;
constexpr int GEMM_THREADS = 64 + 32 * GEMM_EPI_WARPS;
constexpr int GEMM_SLAB_BYTES = 32 * 128;             // 32 rows x 128 bytes
constexpr int GEMM_EPI_WARP_BYTES = 2 * GEMM_SLAB_BYTES + 1024;   // double-buffered slab (+ spare for the token path), 1024-aligned

struct GemmParams {
    int M, N, K;
    void* out;            // bf16 [.., ldc] or fp32 [.., ldc]
    int ldc;
    int out_mode;         // B200X_GEMM_OUT_*
    const float* bias;    // [N] or null
    int act_gelu;
    const float* pe;      // fp32 [group_in, N] (OUT_F32_TOKEN) or null
    int group_in, group_out, group_off;   // out_row = (m / group_in) * group_out + group_off + m % group_in
    long long* prof;      // diagnostic (usually null): per CTA {issuer wait on loads, wait on epilogue, issuer total, tiles}
    int reverse;          // walk the tile list from its end (L2 reuse of the producer kernel's last output, see runtime.cu)
    int a_mmajor;         // single-CTA kernel only: A is [batch][K][128] (M-major: row tile m_blk = batch element, 3-D tensor map)
};

template <int BN>
struct GemmSmem {
    static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
    static constexpr int B_BYTES = BN * GEMM_BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = (GEMM_MAX_STAGES * STAGE_BYTES + GEMM_EPI_WARPS * GEMM_EPI_WARP_BYTES + 1280 <= 232448) ? GEMM_MAX_STAGES : 3;
    static constexpr int EPI_OFFSET = STAGES * STAGE_BYTES;
    static constexpr int EPI_BYTES = GEMM_EPI_WARPS * GEMM_EPI_WARP_BYTES;
    static constexpr int BAR_OFFSET = EPI_OFFSET + EPI_BYTES;
    static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;   // + barriers + alignment slack
    static_assert(B_BYTES % 1024 == 0, "B stage must keep 1024-byte alignment for the 128B swizzle");
    static_assert(TOTAL <= 232448, "shared memory budget exceeded");
};

// Generic (token-mode) chunk: the warp's 32 rows x CW columns sit in `stage` (row stride CW + 4 floats); a group of CW/4
// lanes covers one row with float4s so global accesses are row-contiguous.
template <int CW>
__device__ __forceinline__ void epilogue_chunk_generic(const GemmParams& p, const float* stage, int lane, int m_warp, int n_base) {
    constexpr int LD = CW + 4, LPR = CW / 4, RPI = 32 / LPR, NIT = 32 / RPI;
    const int col = (lane % LPR) * 4;
    const int n0 = n_base + col;
    if (n0 >= p.N) return;
    float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.bias != nullptr) bias4 = *reinterpret_cast<const float4*>(p.bias + n0);
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
        const int row = it * RPI + lane / LPR;
        const int m = m_warp + row;
        if (m >= p.M) continue;
        float4 x = *reinterpret_cast<const float4*>(stage + row * LD + col);
        x.x += bias4.x; x.y += bias4.y; x.z += bias4.z; x.w += bias4.w;
        if (p.act_gelu) { x.x = gelu_fast(x.x); x.y = gelu_fast(x.y); x.z = gelu_fast(x.z); x.w = gelu_fast(x.w); }
        int orow = m, gidx = 0;
        if (p.group_in > 0) {
            gidx = m % p.group_in;
            orow = (m / p.group_in) * p.group_out + p.group_off + gidx;
        }
        if (p.pe != nullptr) {
            const float4 a = *reinterpret_cast<const float4*>(p.pe + static_cast<long long>(gidx) * p.N + n0);
            x.x += a.x; x.y += a.y; x.z += a.z; x.w += a.w;
        }
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + static_cast<long long>(orow) * p.ldc + n0) = x;
    }
}

// acc (+bias) (-> GELU) -> packed bf16 for N consecutive accumulator columns; GELU is a template parameter so that the
// plain path does not carry the (if-converted, predicated-off) GELU instructions through its issue slots
template <int N, bool GELU>
__device__ __forceinline__ void convert_columns(const uint32_t* r, uint32_t* w, const float* bias, int n0, int n_limit) {
#pragma unroll
    for (int i = 0; i < N; i += 4) {
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bias != nullptr && n0 + i < n_limit) b = __ldg(reinterpret_cast<const float4*>(bias + n0 + i));
        float v0 = __uint_as_float(r[i]) + b.x, v1 = __uint_as_float(r[i + 1]) + b.y;
        float v2 = __uint_as_float(r[i + 2]) + b.z, v3 = __uint_as_float(r[i + 3]) + b.w;
        if (GELU) { gelu_fast2(v0, v1); gelu_fast2(v2, v3); }
        w[i / 2] = pack_bf16(v0, v1);
        w[i / 2 + 1] = pack_bf16(v2, v3);
    }
}

// bf16 / residual epilogue of one 128 x BN accumulator tile for one epilogue warp (lane quarter of TMEM, every other
// column group): TMEM -> registers -> bias (+GELU) -> swizzled slab -> one TMA tensor store (or reduce-add) per slab.
template <int BN, int NBUF = 2, int GSTRIDE = 2>
__device__ __forceinline__ void epilogue_store_tile(const GemmParams& p, const CUtensorMap& tmC, const CUtensorMap& tmCtail,
                                                    uint8_t* slab, int& buf, uint32_t t_row, int m_warp, int n_blk, int half, int lane,
                                                    long long* pc = nullptr) {
    const int sw = lane & 7;                             // 128B-swizzle phase of this thread's slab row
    long long pt = 0;
    if (p.out_mode == B200X_GEMM_OUT_BF16) {
        // column groups of 64 (one 128-byte bf16 row per thread); BN = 208 ends with a 16-column group
        constexpr int NG = (BN + 63) / 64;
#pragma unroll 1
        for (int g = half; g < NG; g += GSTRIDE) {
            const int c = g * 64;
            const int n0 = n_blk * BN + c;
            const bool full = (c + 64 <= BN);
            // nothing to store: skip the slab write too - a buffer may only be refilled behind a COMMITTED store, otherwise
            // bulk_wait_read<1> no longer proves that the store which last used it has been read
            if (m_warp >= p.M || n0 >= p.N) continue;
            if (pc) pt = clock64();
            if (NBUF == 2) {
                bulk_wait_read<1>();                     // the store that last used this buffer has read it
                __syncwarp();
            }
            if (pc) { const long long t = clock64(); pc[0] += t - pt; pt = t; }
            uint8_t* dst = slab + buf * GEMM_SLAB_BYTES;
            if (NBUF == 2) buf ^= 1;
            if (full) {
                uint32_t r[64];
                tmem_ld32(t_row + c, r);
                tmem_ld32(t_row + c + 32, r + 32);
                tmem_wait_ld();
                if (pc) { const long long t = clock64(); pc[1] += t - pt; pt = t; }
                uint32_t w[32];
                if (p.act_gelu) convert_columns<64, true>(r, w, p.bias, n0, p.N);
                else convert_columns<64, false>(r, w, p.bias, n0, p.N);
                if (NBUF == 1) {                         // single slab: the previous store had the conversion time to read it
                    bulk_wait_read<0>();
                    __syncwarp();
                }
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<uint4*>(dst + lane * 128 + ((j ^ sw) << 4)) =
                        make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
            } else {
                uint32_t r[16];
                tmem_ld16(t_row + c, r);
                tmem_wait_ld();
                uint32_t w[8];
                if (p.act_gelu) convert_columns<16, true>(r, w, p.bias, n0, p.N);
                else convert_columns<16, false>(r, w, p.bias, n0, p.N);
                if (NBUF == 1) {
                    bulk_wait_read<0>();
                    __syncwarp();
                }
                *reinterpret_cast<uint4*>(dst + lane * 32) = make_uint4(w[0], w[1], w[2], w[3]);       // dense 32-byte rows
                *reinterpret_cast<uint4*>(dst + lane * 32 + 16) = make_uint4(w[4], w[5], w[6], w[7]);
            }
            if (pc) { const long long t = clock64(); pc[2] += t - pt; pt = t; }
            fence_proxy_async();
            __syncwarp();
            if (elect_one()) {
                tma_store_2d(full ? &tmC : &tmCtail, dst, n0, m_warp);   // rows >= M / cols >= N are clipped by TMA
                bulk_commit();
            }
            if (pc) { const long long t = clock64(); pc[3] += t - pt; pt = t; }
        }
    } else if (p.out_mode == B200X_GEMM_OUT_F32_RESID) {
        // column groups of 32 fp32 (128 bytes per thread row); x += acc + bias via TMA reduce-add
        constexpr int NG = BN / 32;
#pragma unroll 1
        for (int g = half; g < NG; g += 2) {
            const int c = g * 32;
            const int n0 = n_blk * BN + c;
            if (m_warp >= p.M || n0 >= p.N) continue;    // see above
            bulk_wait_read<1>();
            __syncwarp();
            uint8_t* dst = slab + buf * GEMM_SLAB_BYTES;
            buf ^= 1;
            uint32_t r[32];
            tmem_ld32(t_row + c, r);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
                if (p.bias != nullptr && n0 + 4 * j < p.N) b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + 4 * j));
                const float4 v = make_float4(__uint_as_float(r[4 * j]) + b.x, __uint_as_float(r[4 * j + 1]) + b.y,
                                             __uint_as_float(r[4 * j + 2]) + b.z, __uint_as_float(r[4 * j + 3]) + b.w);
                *reinterpret_cast<float4*>(dst + lane * 128 + ((j ^ sw) << 4)) = v;
            }
            fence_proxy_async();
            __syncwarp();
            if (elect_one()) {
                tma_reduce_add_2d(&tmC, dst, n0, m_warp);
                bulk_commit();
            }
        }
    }
}

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmCtail, GemmParams p) {
    using L = GemmSmem<BN>;
    constexpr uint32_t TMEM_COLS = (2 * BN <= 256) ? 256 : 512;
    constexpr int STAGES = L::STAGES;
    extern __shared__ uint8_t smem_raw[];
    long long gt_start = 0;
    if (p.prof != nullptr && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_start));
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int m_tiles = (p.M + GEMM_BM - 1) / GEMM_BM;
    const int n_tiles = (p.N + BN - 1) / BN;
    const int num_tiles = m_tiles * n_tiles;
    const int num_kb = (p.K + GEMM_BK - 1) / GEMM_BK;

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
            mbar_init(&tempty_bar[s], GEMM_EPI_WARPS);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (elect_one()) {          // elect.sync: ptxas emits straight-line UTMALDG / UTCHMMA (no per-lane BRA.U.ANY loop)
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * L::STAGE_BYTES;
                    mbar_expect_tx(&full_bar[stage], L::STAGE_BYTES);
                    if (p.a_mmajor) {
                        // M-major A: two [64 k][64 m] boxes (128-byte rows of 64 consecutive m), the m halves 8 KB apart
                        tma_load_3d(sa, &tmA, &full_bar[stage], 0, kb * GEMM_BK, m_blk);
                        tma_load_3d(sa + L::A_BYTES / 2, &tmA, &full_bar[stage], 64, kb * GEMM_BK, m_blk);
                    } else {
                        tma_load_2d(sa, &tmA, &full_bar[stage], kb * GEMM_BK, m_blk * GEMM_BM);
                    }
                    tma_load_2d(sa + L::A_BYTES, &tmB, &full_bar[stage], kb * GEMM_BK, n_blk * BN);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (elect_one()) {          // elect.sync: ptxas emits straight-line UTMALDG / UTCHMMA (no per-lane BRA.U.ANY loop)
            const uint32_t idesc = p.a_mmajor ? make_idesc_bf16(GEMM_BM, BN, false, true) : make_idesc_bf16(GEMM_BM, BN, false);
            int stage = 0;
            uint32_t phase = 0;
            int as = 0;
            uint32_t aphase = 0;
            long long pc_full = 0, pc_te = 0, pc_t = 0, pc_tiles = 0;
            const long long pc_start = p.prof ? clock64() : 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                if (p.prof) pc_t = clock64();
                mbar_wait(&tempty_bar[as], aphase ^ 1);
                tc_fence_after();
                if (p.prof) { pc_te += clock64() - pc_t; ++pc_tiles; }
                const uint32_t d_tmem = tmem_base + as * BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    if (p.prof) pc_t = clock64();
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    if (p.prof) pc_full += clock64() - pc_t;
                    const uint32_t a_addr = smem_u32(smem + stage * L::STAGE_BYTES);
                    const uint32_t b_addr = a_addr + L::A_BYTES;
#pragma unroll
                    for (int k = 0; k < GEMM_BK / 16; ++k) {
                        // K-major A: 16 k = 32 bytes along the 128-byte row; M-major A: 16 k = 16 rows of 128 bytes, the two
                        // 64-wide m blocks 8 KB apart (leading byte offset), 8-row swizzle groups 1 KB apart
                        const uint64_t ad = p.a_mmajor ? make_smem_desc_sw128(a_addr + k * 2048, L::A_BYTES / 2, 1024)
                                                       : make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
                        const uint64_t bd = make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
                        umma_ss(d_tmem, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[stage]);      // frees the smem slot once these MMAs retire
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tfull_bar[as]);             // accumulator ready for the epilogue
                if (++as == 2) { as = 0; aphase ^= 1; }
            }
            if (p.prof) {
                long long* o = p.prof + blockIdx.x * 8;
                o[0] = pc_full; o[1] = pc_te; o[2] = clock64() - pc_start; o[3] = pc_tiles;
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (warps 2..9)
        const int quarter = warp & 3;                    // TMEM lanes [32*quarter, 32*quarter + 32)
        const int half = (warp - 2) >> 2;                // which of the two warps sharing this lane quarter
        uint8_t* slab = smem + L::EPI_OFFSET + (warp - 2) * GEMM_EPI_WARP_BYTES;
        int as = 0, buf = 0, tile_parity = 0;
        uint32_t aphase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_parity) {
            const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
            mbar_wait(&tfull_bar[as], aphase);
            tc_fence_after();
            const int m_warp = m_blk * GEMM_BM + quarter * 32;
            const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * BN;
            if (p.out_mode != B200X_GEMM_OUT_F32_TOKEN) {
                epilogue_store_tile<BN>(p, tmC, tmCtail, slab, buf, t_row, m_warp, n_blk, half ^ (tile_parity & 1), lane);
            } else {
                // token mode: generic transposing path (row remap + positional encoding), 32-column chunks
                float* stage = reinterpret_cast<float*>(slab);
                bulk_wait_read<0>();
                __syncwarp();
                constexpr int NCHUNK = (BN + 31) / 32;
#pragma unroll 1
                for (int ci = half; ci < NCHUNK; ci += 2) {
                    const int c = ci * 32;
                    const bool wide = (c + 32 <= BN);
                    uint32_t r[32];
                    if (wide) tmem_ld32(t_row + c, r); else tmem_ld16(t_row + c, r);
                    tmem_wait_ld();
                    if (wide) {
#pragma unroll
                        for (int i = 0; i < 32; i += 4)
                            *reinterpret_cast<uint4*>(stage + lane * 36 + i) = make_uint4(r[i], r[i + 1], r[i + 2], r[i + 3]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; i += 4)
                            *reinterpret_cast<uint4*>(stage + lane * 20 + i) = make_uint4(r[i], r[i + 1], r[i + 2], r[i + 3]);
                    }
                    __syncwarp();
                    if (wide) epilogue_chunk_generic<32>(p, stage, lane, m_warp, n_blk * BN + c);
                    else epilogue_chunk_generic<16>(p, stage, lane, m_warp, n_blk * BN + c);
                    __syncwarp();
                }
            }
            tc_fence_before();
            __syncwarp();
            if (elect_one()) mbar_arrive(&tempty_bar[as]);
            if (++as == 2) { as = 0; aphase ^= 1; }
        }
        bulk_wait<0>();                   // all tensor stores of this warp have completed
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
    if (p.prof != nullptr && threadIdx.x == 0) {
        long long gt_end;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_end));
        p.prof[blockIdx.x * 8 + 4] = gt_start;
        p.prof[blockIdx.x * 8 + 5] = gt_end;
    }
}

// ------------------------------------------------------------------------------------------------ CTA-pair GEMM
// Two CTAs of a cluster (one TPC) share one 256 x BN tile through tcgen05 cta_group::2: each CTA loads its own 128 rows
// of A and HALF of the B tile (BN/2 weight rows) and the pair's tensor cores exchange the halves.  A single-CTA
// 128 x 192 tile has to pull 40 KB through the SM's L2 port (~54 B/clk measured) per 64-deep k-block of 424 MMA
// cycles - the port, not the tensor pipe, bounds it (tools/gemm_profile.py).  The pair halves the B bytes per SM.
//   warp 0 (both CTAs): TMA producer for its own shared memory; completion bytes go to the LEADER's full barrier
//   warp 1 (leader)   : issues tcgen05.mma.cta_group::2; commits are multicast to both CTAs' empty / tfull barriers
//   warps 2-9 (both)  : epilogue of the CTA's own 128 accumulator rows; "accumulator drained" arrives on the leader
template <int BN>
struct Gemm2Smem {
    static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
    static constexpr int B_BYTES = (BN / 2) * GEMM_BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = (5 * STAGE_BYTES + GEMM_EPI_WARPS * GEMM_EPI_WARP_BYTES + 1280 <= 232448) ? 5 : 4;
    static constexpr int EPI_OFFSET = STAGES * STAGE_BYTES;
    static constexpr int BAR_OFFSET = EPI_OFFSET + GEMM_EPI_WARPS * GEMM_EPI_WARP_BYTES;
    static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;
    static_assert(B_BYTES % 1024 == 0, "B half tile must keep 1024-byte alignment for the 128B swizzle");
    static_assert(TOTAL <= 232448, "shared memory budget exceeded");
};

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm2_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmCtail, GemmParams p) {
    using L = Gemm2Smem<BN>;
    constexpr uint32_t TMEM_COLS = (2 * BN <= 256) ? 256 : 512;
    constexpr int STAGES = L::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int m_tiles = (p.M + 2 * GEMM_BM - 1) / (2 * GEMM_BM);
    const int n_tiles = (p.N + BN - 1) / BN;
    const int num_tiles = m_tiles * n_tiles;
    const int num_kb = (p.K + GEMM_BK - 1) / GEMM_BK;

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
            mbar_init(&tempty_bar[s], 2 * GEMM_EPI_WARPS);      // the epilogue warps of BOTH CTAs (leader's copy is used)
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_pair<TMEM_COLS>(tmem_slot);
    tc_fence_before();
    cluster_sync_all();                                  // barrier inits and the TMEM allocation are visible pair-wide
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer (both CTAs)
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = pair; tile < num_tiles; tile += n_pairs) {
                const int tile_e = p.reverse ? num_tiles - 1 - tile : tile;
                const int m_blk = tile_e / n_tiles, n_blk = tile_e % n_tiles;
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
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (leader CTA only)
        if (rank == 0 && elect_one()) {
            constexpr uint32_t idesc = make_idesc_bf16(2 * GEMM_BM, BN, false);
            int stage = 0;
            uint32_t phase = 0;
            int as = 0;
            uint32_t aphase = 0;
            long long pc_full = 0, pc_te = 0, pc_t = 0, pc_tiles = 0;
            const long long pc_start = p.prof ? clock64() : 0;
            for (int tile = pair; tile < num_tiles; tile += n_pairs) {
                if (p.prof) pc_t = clock64();
                mbar_wait(&tempty_bar[as], aphase ^ 1);
                tc_fence_after();
                if (p.prof) { pc_te += clock64() - pc_t; ++pc_tiles; }
                const uint32_t d_tmem = tmem_base + as * BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    if (p.prof) pc_t = clock64();
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    if (p.prof) pc_full += clock64() - pc_t;
                    const uint32_t a_addr = smem_u32(smem + stage * L::STAGE_BYTES);
                    const uint32_t b_addr = a_addr + L::A_BYTES;
#pragma unroll
                    for (int k = 0; k < GEMM_BK / 16; ++k) {
                        const uint64_t ad = make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
                        const uint64_t bd = make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
                        umma_ss_pair(d_tmem, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit_pair(&empty_bar[stage]);  // frees the slot in both CTAs once these MMAs retire
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit_pair(&tfull_bar[as]);         // accumulator ready for both CTAs' epilogues
                if (++as == 2) { as = 0; aphase ^= 1; }
            }
            if (p.prof) {
                long long* o = p.prof + blockIdx.x * 8;
                o[0] = pc_full; o[1] = pc_te; o[2] = clock64() - pc_start; o[3] = pc_tiles;
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (warps 2..9 of both CTAs)
        const int quarter = warp & 3;
        const int half = (warp - 2) >> 2;
        uint8_t* slab = smem + L::EPI_OFFSET + (warp - 2) * GEMM_EPI_WARP_BYTES;
        int as = 0, buf = 0, tile_parity = 0;
        uint32_t aphase = 0;
        const uint32_t leader_tempty0 = mapa_shared(&tempty_bar[0], 0), leader_tempty1 = mapa_shared(&tempty_bar[1], 0);
        long long pce[4] = {0, 0, 0, 0}, pc_wt = 0, pc_t0 = 0;
        long long* pc = (p.prof != nullptr && warp == 2 && rank == 0) ? pce : nullptr;
        for (int tile = pair; tile < num_tiles; tile += n_pairs, ++tile_parity) {
            const int tile_e = p.reverse ? num_tiles - 1 - tile : tile;
            const int m_blk = tile_e / n_tiles, n_blk = tile_e % n_tiles;
            if (pc) pc_t0 = clock64();
            mbar_wait(&tfull_bar[as], aphase);
            tc_fence_after();
            if (pc) pc_wt += clock64() - pc_t0;
            const int m_warp = m_blk * 2 * GEMM_BM + static_cast<int>(rank) * GEMM_BM + quarter * 32;
            const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * BN;
            epilogue_store_tile<BN>(p, tmC, tmCtail, slab, buf, t_row, m_warp, n_blk, half ^ (tile_parity & 1), lane, pc);
            tc_fence_before();
            __syncwarp();
            if (elect_one()) mbar_arrive_cluster(as == 0 ? leader_tempty0 : leader_tempty1);
            if (++as == 2) { as = 0; aphase ^= 1; }
        }
        if (pc != nullptr && lane == 0) {
            long long* o = p.prof + blockIdx.x * 8 + 6;
            o[0] = (pce[0] << 32) | (pce[1] & 0xffffffffll);
            o[1] = (pce[2] << 32) | (pce[3] & 0xffffffffll);
            p.prof[(blockIdx.x + 1) * 8 + 6] = pc_wt;
        }
        bulk_wait<0>();                                  // all tensor stores of this warp have completed
    }

    tc_fence_before();
    cluster_sync_all();                                  // no CTA of the pair may free TMEM / exit while the other still uses it
    if (warp == 1) tmem_dealloc_pair<TMEM_COLS>(tmem_base);
}


// ------------------------------------------------------------------------------------------------ A-stationary CTA-pair GEMM
// For wide outputs of a narrow K (QKV: N = 1152 = 6 column tiles, K = 384): the pair kernel above re-reads the 128 x K row tile
// of A from L2 once per column tile, and the L2 -> SM operand traffic is what paces it.  Here a CTA keeps its row tile in
// shared memory (K <= 384: six 16 KB k-blocks, loaded ONCE per row tile by a third single-thread role, k-block by k-block as
// the last column tile of the previous row tile releases them) and sweeps all column tiles of W over it:
//   * a stage of the ring is only the B half tile (12 KB), so the ring is 7-8 stages deep and covers the ~3 k-cycle L2 -> SM
//     round trip (in-kernel counters: with 4 stages the issuer waited 390 cycles per k-step for weights);
//   * the pairs walk the column tiles in ROTATED orders - in lockstep all 74 pairs pull the same 24 KB weight tile out of the
//     same few L2 slices at the same time;
//   * the epilogue warps make do with one output slab each.
// Same MMAs in the same order as gemm2_bf16_tn_kernel: results are bit-identical.  Measured at 229 copies (profiles/
// r02_g_gemm_ln_astationary.txt, debug = 1 rows): 252-265 us against 267-270 us (1.05-1.11 PFLOP/s).
constexpr int B200X_FC1_EPI_WARPS = 12;   // fc1 (GELU epilogue): 271.5-272.4 us with 12 epilogue warps against 276.3-278.2 us with 8 (profiles/r02_x_fc1_epilogue_warps.txt)
constexpr int GEMM_AS_MAX_KB = 6;
// EW = epilogue warps: 8 (two per TMEM lane quarter) or 12 (three per quarter, for the GELU epilogue of fc1, which is
// bound by the epilogue's dependent chains: TMEM load -> erf-GELU -> pack -> slab -> TMA store)
template <int EW> struct GemmAsRoles {
    static constexpr int W_TMA_A = 2 + EW;
    static constexpr int THREADS = 32 * (W_TMA_A + 1);
};

template <int BN, int EW>
struct GemmAsSmem {
    static constexpr int A_KB_BYTES = GEMM_BM * GEMM_BK * 2;                  // 16 KB per k-block
    static constexpr int A_BYTES = GEMM_AS_MAX_KB * A_KB_BYTES;               // 96 KB
    static constexpr int B_BYTES = (BN / 2) * GEMM_BK * 2;
    static constexpr int EPI_WARP_BYTES = GEMM_SLAB_BYTES;
    static constexpr int STAGES_FIT = (232448 - 1024 - 512 - A_BYTES - EW * EPI_WARP_BYTES) / B_BYTES;
    static constexpr int STAGES = STAGES_FIT > 8 ? 8 : STAGES_FIT;
    static constexpr int B_OFFSET = A_BYTES;
    static constexpr int EPI_OFFSET = B_OFFSET + STAGES * B_BYTES;
    static constexpr int BAR_OFFSET = EPI_OFFSET + EW * EPI_WARP_BYTES;
    static constexpr int TOTAL = BAR_OFFSET + 512 + 1024;
    static_assert(B_BYTES % 1024 == 0, "B half tile must keep 1024-byte alignment for the 128B swizzle");
    static_assert(STAGES >= 4 && TOTAL <= 232448, "shared memory budget exceeded");
};

template <int BN, int EW>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GemmAsRoles<EW>::THREADS, 1)
gemm2_astat_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                           const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmCtail, GemmParams p) {
    using L = GemmAsSmem<BN, EW>;
    constexpr int SUBS = EW / 4;                         // epilogue warps per TMEM lane quarter
    constexpr uint32_t TMEM_COLS = (2 * BN <= 256) ? 256 : 512;
    constexpr int STAGES = L::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint64_t* a_full = tempty_bar + 2;                   // [6] k-block kb of the row tile has landed in BOTH CTAs (leader's copy)
    uint64_t* a_empty = a_full + GEMM_AS_MAX_KB;         // [6] per CTA: the row tile's last MMAs on k-block kb have retired
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_empty + GEMM_AS_MAX_KB);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int m_tiles = (p.M + 2 * GEMM_BM - 1) / (2 * GEMM_BM);
    const int n_tiles = (p.N + BN - 1) / BN;
    const int num_kb = (p.K + GEMM_BK - 1) / GEMM_BK;
    const int n_rot = pair % n_tiles;

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
            mbar_init(&tempty_bar[s], 2 * EW);
        }
        for (int s = 0; s < GEMM_AS_MAX_KB; ++s) {
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
            for (int mt = pair; mt < m_tiles; mt += n_pairs, tphase ^= 1) {
                for (int j = 0; j < n_tiles; ++j) {
                    mbar_wait(&tempty_bar[as], aphase ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + as * BN;
                    for (int kb = 0; kb < num_kb; ++kb) {
                        if (j == 0) mbar_wait(&a_full[kb], tphase);              // both CTAs' k-block has landed
                        mbar_wait(&full_bar[stage], phase);
                        tc_fence_after();
                        const uint32_t a_addr = smem_u32(smem + kb * L::A_KB_BYTES);
                        const uint32_t b_addr = smem_u32(smem + L::B_OFFSET + stage * L::B_BYTES);
#pragma unroll
                        for (int k = 0; k < GEMM_BK / 16; ++k) {
                            const uint64_t ad = make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
                            const uint64_t bd = make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
                            umma_ss_pair(d_tmem, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
                        }
                        umma_commit_pair(&empty_bar[stage]);
                        if (j == n_tiles - 1) umma_commit_pair(&a_empty[kb]);    // last reader of this k-block: refill it
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                    umma_commit_pair(&tfull_bar[as]);
                    if (++as == 2) { as = 0; aphase ^= 1; }
                }
            }
        }
    } else if (warp < GemmAsRoles<EW>::W_TMA_A) {
        // ------------------------------------------------------------------ epilogue (warps 2..2+EW-1 of both CTAs)
        const int quarter = warp & 3;
        const int sub = (warp - 2) >> 2;
        uint8_t* slab = smem + L::EPI_OFFSET + (warp - 2) * L::EPI_WARP_BYTES;
        int as = 0, buf = 0, tile_parity = 0;
        uint32_t aphase = 0;
        const uint32_t leader_tempty0 = mapa_shared(&tempty_bar[0], 0), leader_tempty1 = mapa_shared(&tempty_bar[1], 0);
        for (int mt = pair; mt < m_tiles; mt += n_pairs) {
            const int m_blk = p.reverse ? m_tiles - 1 - mt : mt;
            for (int j = 0; j < n_tiles; ++j, ++tile_parity) {
                const int n_blk = (j + n_rot) % n_tiles;
                mbar_wait(&tfull_bar[as], aphase);
                tc_fence_after();
                const int m_warp = m_blk * 2 * GEMM_BM + static_cast<int>(rank) * GEMM_BM + quarter * 32;
                const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * BN;
                epilogue_store_tile<BN, 1, SUBS>(p, tmC, tmCtail, slab, buf, t_row, m_warp, n_blk, (sub + tile_parity) % SUBS, lane);
                tc_fence_before();
                __syncwarp();
                if (elect_one()) mbar_arrive_cluster(as == 0 ? leader_tempty0 : leader_tempty1);
                if (++as == 2) { as = 0; aphase ^= 1; }
            }
        }
        bulk_wait<0>();
    } else {
        // ------------------------------------------------------------------ TMA-A: the stationary row tile, k-block by k-block
        if (elect_one()) {
            int it = 0;
            for (int mt = pair; mt < m_tiles; mt += n_pairs, ++it) {
                const int m_blk = p.reverse ? m_tiles - 1 - mt : mt;
                const int row = m_blk * 2 * GEMM_BM + static_cast<int>(rank) * GEMM_BM;
                for (int kb = 0; kb < num_kb; ++kb) {
                    if (it > 0) mbar_wait(&a_empty[kb], (it - 1) & 1);       // the previous row tile no longer reads this k-block
                    if (rank == 0) mbar_expect_tx(&a_full[kb], 2 * L::A_KB_BYTES);
                    tma_load_2d_pair(smem + kb * L::A_KB_BYTES, &tmA, mapa_shared(&a_full[kb], 0), kb * GEMM_BK, row);
                }
            }
        }
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) tmem_dealloc_pair<TMEM_COLS>(tmem_base);
}

template <int BN, int EW>
static int launch_gemm2_astat(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmCtail,
                              const GemmParams& p, cudaStream_t stream) {
    using L = GemmAsSmem<BN, EW>;
    int num_sms = 0;
    B200X_TRY(ensure_kernel_smem(reinterpret_cast<const void*>(gemm2_astat_bf16_tn_kernel<BN, EW>), L::TOTAL));
    B200X_TRY(device_sm_count(&num_sms));
    const int pairs = std::min(ceil_div(p.M, 2 * GEMM_BM), num_sms / 2);
    gemm2_astat_bf16_tn_kernel<BN, EW><<<2 * pairs, GemmAsRoles<EW>::THREADS, L::TOTAL, stream>>>(tmA, tmB, tmC, tmCtail, p);
    B200X_CUDA_TRY(cudaGetLastError());
    return B200X_OK;
}

// ------------------------------------------------------------------------------------------------ residual GEMM + LayerNorm tail
// x += A . W^T + bias (fp32 residual stream, TMA reduce-add as above), then h = LayerNorm(x) in bf16 for the projection that
// follows (attention proj -> LN2 -> fc1; fc2 -> next block's LN1 -> QKV).  The separate LayerNorm pass read the 484 MB
// residual stream back from HBM a few microseconds after the residual GEMM had updated it; here a CTA normalises its own 128
// rows as soon as its reduce-adds for ALL column tiles of the row tile have completed, while they still sit in L2:
//   warps 16 / 17: TMA producer / MMA issuer as in gemm2_bf16_tn_kernel, but row-tile major: the pair finishes every column
//                  tile of a 256-row tile before it moves on
//   warps 8-15   : epilogue (reduce-add); after the last column tile each warp waits for its bulk groups to COMPLETE
//                  (cp.async.bulk.wait_group 0: the adds are performed in L2) and arrives on x_ready[row tile & 1]
//   warps 0-7    : LayerNorm tail, 16 rows per warp, one warp per row, the arithmetic of layernorm_kernel bit for bit (same
//                  summation tree), four rows in flight per warp; ld.global.cg (L2) -> bf16 h rows.  They lag one row tile
//                  behind the MMAs; ln_done[] stops the epilogue from lapping them.
// Measured (229 copies, profiles/r02_h_gemm_resid_ln.txt): fc2 + tail 379 us against 310 + 107 us for the two kernels, but
// proj + tail 322 us against 211 + 107 us: ncu shows the tail's re-read of x coming from DRAM after all (1.17 GB read per call
// = A + reduce-add RMW + 0.45 GB) - lines produced by TMA reduce-adds are not served from L2 to later readers (an evict_last
// hint on the reduction, or prefetching the lines ahead of it, changes nothing).  Behind the K = 384 projection the kernel
// therefore runs its load-add-store epilogue (LS, below): 303 us, 0.99 GB read.
constexpr int GEMM_RLN_LN_WARPS = 8;                  // 4 warps: fc2 + tail 469 us, 16 warps (72 registers): 385-417 us, 8: 377 us
                                                      // (profiles/r02_h_gemm_resid_ln.txt, last block)
// warp roles: the SM's warp arbiter prefers HIGHER warp ids, and a TMA producer / MMA issuer that loses its issue slots to the
// eight busy LayerNorm warps delays every tcgen05.mma - so the single-thread roles sit on top, the LayerNorm tail at the bottom
constexpr int GEMM_RLN_W_EPI = GEMM_RLN_LN_WARPS;                  // 8..15 (quarter = warp & 3 still names the TMEM lane quarter)
constexpr int GEMM_RLN_W_TMA = GEMM_RLN_W_EPI + GEMM_EPI_WARPS;   // 16
constexpr int GEMM_RLN_W_MMA = GEMM_RLN_W_TMA + 1;                // 17
constexpr int GEMM_RLN_THREADS = 32 * (GEMM_RLN_W_MMA + 1);
constexpr int GEMM_RLN_ROWS_PER_BATCH = 4;

struct GemmLnTail {
    const float* x;          // [M][ldx] fp32 residual stream (the GEMM's reduce-add target)
    int ldx;
    const float* gamma;      // [N]
    const float* beta;       // [N]
    float eps;
    __nv_bfloat16* h;        // [M][ldh] bf16 LayerNorm(x)
    int ldh;
};

// fp32 residual epilogue of one accumulator tile for one epilogue warp: x += acc + bias via TMA reduce-add
template <int BN>
__device__ __forceinline__ void epilogue_resid_tile(const GemmParams& p, const CUtensorMap& tmC, uint8_t* slab, int& buf, uint32_t t_row,
                                                    int m_warp, int n_blk, int half, int lane) {
    const int sw = lane & 7;
    constexpr int NG = BN / 32;
#pragma unroll 1
    for (int g = half; g < NG; g += 2) {
        const int c = g * 32;
        const int n0 = n_blk * BN + c;
        if (m_warp >= p.M || n0 >= p.N) continue;
        bulk_wait_read<1>();
        __syncwarp();
        uint8_t* dst = slab + buf * GEMM_SLAB_BYTES;
        buf ^= 1;
        uint32_t r[32];
        tmem_ld32(t_row + c, r);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.bias != nullptr && n0 + 4 * j < p.N) b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + 4 * j));
            const float4 v = make_float4(__uint_as_float(r[4 * j]) + b.x, __uint_as_float(r[4 * j + 1]) + b.y,
                                         __uint_as_float(r[4 * j + 2]) + b.z, __uint_as_float(r[4 * j + 3]) + b.w);
            *reinterpret_cast<float4*>(dst + lane * 128 + ((j ^ sw) << 4)) = v;
        }
        fence_proxy_async();
        __syncwarp();
        if (elect_one()) {
            tma_reduce_add_2d(&tmC, dst, n0, m_warp);
            bulk_commit();
        }
    }
}

// Load-add-store form of the residual epilogue (LS = true): x_old comes in by TMA (32 x 32 fp32 boxes, requested one tile ahead),
// v = (acc + bias) + x_old is formed in registers - the same sum the reduce-add forms in L2, bit for bit - and goes back by a plain
// TMA store from the slab it arrived in.  Plain stores leave their lines in L2, so the LayerNorm tail's re-read of the row tile
// is served from L2; the result of a TMA reduce-add is not (the tail behind the reduce-add epilogue re-read 0.45 GB per call from
// DRAM).  Three 4 KB slabs per epilogue warp (one per column group of a tile) leave room for a four-stage ring.
constexpr int GEMM_RLS_SLABS = 3;
template <int BN>
struct GemmRlsSmem {
    static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
    static constexpr int B_BYTES = (BN / 2) * GEMM_BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int EPI_WARP_BYTES = GEMM_RLS_SLABS * GEMM_SLAB_BYTES;
    static constexpr int STAGES = (232448 - 1024 - 512 - GEMM_EPI_WARPS * EPI_WARP_BYTES) / STAGE_BYTES;
    static constexpr int EPI_OFFSET = STAGES * STAGE_BYTES;
    static constexpr int BAR_OFFSET = EPI_OFFSET + GEMM_EPI_WARPS * EPI_WARP_BYTES;
    static constexpr int TOTAL = BAR_OFFSET + 512 + 1024;
    static_assert(BN == 192, "three column groups per epilogue warp and tile");
    static_assert(STAGES >= 4 && TOTAL <= 232448, "shared memory budget exceeded");
};

template <int BN, bool LS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_RLN_THREADS, 1)
gemm2_resid_ln_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ CUtensorMap tmC, GemmParams p, GemmLnTail q) {
    using L = typename std::conditional<LS, GemmRlsSmem<BN>, Gemm2Smem<BN>>::type;
    constexpr uint32_t TMEM_COLS = (2 * BN <= 256) ? 256 : 512;
    constexpr int STAGES = L::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint64_t* x_ready = tempty_bar + 2;                  // [2] this CTA's rows of the row tile are final in L2
    uint64_t* ln_done = x_ready + 2;                     // [2] the LayerNorm warps have consumed x_ready[b]
    uint64_t* xfull_bar = ln_done + 2;                   // LS: [8 warps][3] x_old box has landed in the warp's slab
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xfull_bar + (LS ? GEMM_EPI_WARPS * GEMM_RLS_SLABS : 0));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int m_tiles = (p.M + 2 * GEMM_BM - 1) / (2 * GEMM_BM);
    const int n_tiles = (p.N + BN - 1) / BN;
    const int num_kb = (p.K + GEMM_BK - 1) / GEMM_BK;

    if (warp == GEMM_RLN_W_TMA && elect_one()) {
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
            mbar_init(&x_ready[s], GEMM_EPI_WARPS);
            mbar_init(&ln_done[s], GEMM_RLN_LN_WARPS);
        }
        if (LS) for (int s = 0; s < GEMM_EPI_WARPS * GEMM_RLS_SLABS; ++s) mbar_init(&xfull_bar[s], 1);
        fence_barrier_init();
    }
    if (warp == GEMM_RLN_W_MMA) tmem_alloc_pair<TMEM_COLS>(tmem_slot);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == GEMM_RLN_W_TMA) {
        // ------------------------------------------------------------------ TMA producer (both CTAs)
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int mt = pair; mt < m_tiles; mt += n_pairs) {
                const int m_blk = p.reverse ? m_tiles - 1 - mt : mt;
                for (int n_blk = 0; n_blk < n_tiles; ++n_blk) {
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
    } else if (warp == GEMM_RLN_W_MMA) {
        // ------------------------------------------------------------------ MMA issuer (leader CTA only)
        if (rank == 0 && elect_one()) {
            constexpr uint32_t idesc = make_idesc_bf16(2 * GEMM_BM, BN, false);
            int stage = 0, as = 0;
            uint32_t phase = 0, aphase = 0;
            for (int mt = pair; mt < m_tiles; mt += n_pairs) {
                for (int n_blk = 0; n_blk < n_tiles; ++n_blk) {
                    mbar_wait(&tempty_bar[as], aphase ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + as * BN;
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
                    umma_commit_pair(&tfull_bar[as]);
                    if (++as == 2) { as = 0; aphase ^= 1; }
                }
            }
        }
    } else if (warp >= GEMM_RLN_W_EPI) {
        // ------------------------------------------------------------------ epilogue (warps 8..15 of both CTAs)
        const int quarter = warp & 3;
        const int half = (warp - GEMM_RLN_W_EPI) >> 2;
        uint8_t* slab = smem + L::EPI_OFFSET + (warp - GEMM_RLN_W_EPI) * (LS ? GEMM_RLS_SLABS * GEMM_SLAB_BYTES : GEMM_EPI_WARP_BYTES);
        int as = 0, buf = 0, tile_parity = 0, it = 0;
        uint32_t aphase = 0;
        const uint32_t leader_tempty0 = mapa_shared(&tempty_bar[0], 0), leader_tempty1 = mapa_shared(&tempty_bar[1], 0);
        // LS: the x_old box of column group li (0..2) of tile `seq` of this warp lives in slab li; seq counts this pair's tiles
        uint64_t* xfull = xfull_bar + (warp - GEMM_RLN_W_EPI) * GEMM_RLS_SLABS;
        uint32_t xph = 0;                                // bit li: parity of the next completion of xfull[li]
        const uint64_t pol_last = l2_policy_evict_last();   // the tail re-reads these lines within a tile time
        const int my_tiles = (m_tiles - pair + n_pairs - 1) / n_pairs * n_tiles;
        auto tile_of = [&](int seq, int& m_warp_o, int& n_blk_o, int& half_o) {
            const int mt_s = pair + (seq / n_tiles) * n_pairs;
            const int m_blk_s = p.reverse ? m_tiles - 1 - mt_s : mt_s;
            m_warp_o = m_blk_s * 2 * GEMM_BM + static_cast<int>(rank) * GEMM_BM + quarter * 32;
            n_blk_o = seq % n_tiles;
            half_o = half ^ (seq & 1);
        };
        auto load_x = [&](int seq, int li) {             // elected lane: request x_old of group li of tile seq (if it exists)
            int mw, nb, hf;
            tile_of(seq, mw, nb, hf);
            const int n0 = nb * BN + (hf + 2 * li) * 32;
            if (mw >= p.M || n0 >= p.N) return;
            mbar_expect_tx(&xfull[li], GEMM_SLAB_BYTES);
            tma_load_2d(slab + li * GEMM_SLAB_BYTES, &tmC, &xfull[li], n0, mw);
        };
        if (LS && my_tiles > 0 && elect_one()) {
#pragma unroll
            for (int li = 0; li < GEMM_RLS_SLABS; ++li) load_x(0, li);
        }
        int seq = 0;
        for (int mt = pair; mt < m_tiles; mt += n_pairs, ++it) {
            const int m_blk = p.reverse ? m_tiles - 1 - mt : mt;
            for (int n_blk = 0; n_blk < n_tiles; ++n_blk, ++tile_parity, ++seq) {
                mbar_wait(&tfull_bar[as], aphase);
                tc_fence_after();
                const int m_warp = m_blk * 2 * GEMM_BM + static_cast<int>(rank) * GEMM_BM + quarter * 32;
                const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * BN;
                if (LS) {
                    const int hf = half ^ (seq & 1);
                    const int sw = lane & 7;
                    const bool more = seq + 1 < my_tiles;
#pragma unroll 1
                    for (int li = 0; li < GEMM_RLS_SLABS; ++li) {
                        const int c = (hf + 2 * li) * 32;
                        const int n0 = n_blk * BN + c;
                        if (m_warp >= p.M || n0 >= p.N) continue;        // no load was requested for it either (load_x)
                        uint8_t* xs = slab + li * GEMM_SLAB_BYTES;
                        uint32_t r[32];
                        tmem_ld32(t_row + c, r);
                        mbar_wait(&xfull[li], (xph >> li) & 1);
                        xph ^= 1u << li;
                        tmem_wait_ld();
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (p.bias != nullptr && n0 + 4 * j < p.N) b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + 4 * j));
                            float4* cell = reinterpret_cast<float4*>(xs + lane * 128 + ((j ^ sw) << 4));
                            const float4 xo = *cell;
                            *cell = make_float4((__uint_as_float(r[4 * j]) + b.x) + xo.x, (__uint_as_float(r[4 * j + 1]) + b.y) + xo.y,
                                                (__uint_as_float(r[4 * j + 2]) + b.z) + xo.z, (__uint_as_float(r[4 * j + 3]) + b.w) + xo.w);
                        }
                        fence_proxy_async();
                        __syncwarp();
                        if (elect_one()) {
                            tma_store_2d_hint(&tmC, xs, n0, m_warp, pol_last);   // rows >= M are clipped by TMA; evict_last: -2 %
                            bulk_commit();
                        }
                    }
                    // the next tile's x_old boxes, requested a whole tile ahead (independently of what this tile skipped:
                    // with a ragged last row tile walked first, a warp without rows here still has rows in the next tile)
                    if (more && elect_one()) {
                        bulk_wait_read<0>();                             // the slabs' stores have been read out
#pragma unroll
                        for (int li = 0; li < GEMM_RLS_SLABS; ++li) load_x(seq + 1, li);
                    }
                } else {
                    epilogue_resid_tile<BN>(p, tmC, slab, buf, t_row, m_warp, n_blk, half ^ (tile_parity & 1), lane);
                }
                tc_fence_before();
                __syncwarp();
                if (elect_one()) mbar_arrive_cluster(as == 0 ? leader_tempty0 : leader_tempty1);
                if (++as == 2) { as = 0; aphase ^= 1; }
            }
            // every reduce-add of this warp into the row tile has been PERFORMED (not merely read out of the slab)
            bulk_wait<0>();
            const int b = it & 1, u = it >> 1;
            if (u > 0) mbar_wait(&ln_done[b], (u - 1) & 1);   // the LayerNorm warps have seen the previous use of x_ready[b]
            __syncwarp();
            if (elect_one()) mbar_arrive(&x_ready[b]);
        }
    } else {
        // ------------------------------------------------------------------ LayerNorm tail (warps 0..7 of both CTAs)
        constexpr int ROWS_PER_WARP = GEMM_BM / GEMM_RLN_LN_WARPS;           // 16
        constexpr int RB = GEMM_RLN_ROWS_PER_BATCH;
        const int pw = warp;
        const int D = p.N;
        const int vpl = D / 128;                                             // float4 vectors per lane (<= 3)
        const float inv_d = 1.0f / static_cast<float>(D);
        int it = 0;
        for (int mt = pair; mt < m_tiles; mt += n_pairs, ++it) {
            const int m_blk = p.reverse ? m_tiles - 1 - mt : mt;
            const int b = it & 1, u = it >> 1;
            mbar_wait(&x_ready[b], u & 1);
            fence_proxy_async_all();                                         // async-proxy (TMA reduce) writes -> generic-proxy reads
            const long long row0 = static_cast<long long>(m_blk) * 2 * GEMM_BM + static_cast<long long>(rank) * GEMM_BM + pw * ROWS_PER_WARP;
#pragma unroll 1
            for (int r0 = 0; r0 < ROWS_PER_WARP; r0 += RB) {
                if (row0 + r0 >= p.M) break;
                // RB rows in flight per warp; every step runs over all RB rows before the next one starts, so the shuffle /
                // divide / rsqrt latencies of the rows overlap (written row after row, ptxas serialises the chains)
                float4 v[RB][3];
#pragma unroll
                for (int r = 0; r < RB; ++r) {
                    const long long row = row0 + r0 + r;
                    const float4* src = reinterpret_cast<const float4*>(q.x + row * q.ldx);
#pragma unroll
                    for (int i = 0; i < 3; ++i)
                        v[r][i] = (row < p.M && i < vpl) ? __ldcg(src + lane + 32 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                layernorm_rows<RB>(v, vpl, inv_d, q.eps, q.gamma, q.beta, lane, [&](int r, int i, float4 o) {
                    const long long row = row0 + r0 + r;
                    if (row < p.M)
                        reinterpret_cast<uint2*>(q.h + row * q.ldh)[lane + 32 * i] = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
                });
            }
            __syncwarp();
            if (elect_one()) mbar_arrive(&ln_done[b]);
        }
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == GEMM_RLN_W_MMA) tmem_dealloc_pair<TMEM_COLS>(tmem_base);
}

template <int BN, bool LS>
static int launch_gemm2_resid_ln(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const GemmParams& p,
                                 const GemmLnTail& q, cudaStream_t stream) {
    using L = typename std::conditional<LS, GemmRlsSmem<BN>, Gemm2Smem<BN>>::type;
    int num_sms = 0;
    B200X_TRY(ensure_kernel_smem(reinterpret_cast<const void*>(gemm2_resid_ln_kernel<BN, LS>), L::TOTAL));
    B200X_TRY(device_sm_count(&num_sms));
    const int pairs = std::min(ceil_div(p.M, 2 * GEMM_BM), num_sms / 2);
    gemm2_resid_ln_kernel<BN, LS><<<2 * pairs, GEMM_RLN_THREADS, L::TOTAL, stream>>>(tmA, tmB, tmC, p, q);
    B200X_CUDA_TRY(cudaGetLastError());
    return B200X_OK;
}

template <int BN>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmCtail,
                       const GemmParams& p, cudaStream_t stream) {
    using L = GemmSmem<BN>;
    int num_sms = 0;
    B200X_TRY(ensure_kernel_smem(reinterpret_cast<const void*>(gemm_bf16_tn_kernel<BN>), L::TOTAL));
    B200X_TRY(device_sm_count(&num_sms));
    const int tiles = ceil_div(p.M, GEMM_BM) * ceil_div(p.N, BN);
    const int grid = tiles < num_sms ? tiles : num_sms;
    gemm_bf16_tn_kernel<BN><<<grid, GEMM_THREADS, L::TOTAL, stream>>>(tmA, tmB, tmC, tmCtail, p);
    B200X_CUDA_TRY(cudaGetLastError());
    return B200X_OK;
}

template <int BN>
static int launch_gemm2(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmCtail,
                        const GemmParams& p, cudaStream_t stream) {
    using L = Gemm2Smem<BN>;
    int num_sms = 0;
    B200X_TRY(ensure_kernel_smem(reinterpret_cast<const void*>(gemm2_bf16_tn_kernel<BN>), L::TOTAL));
    B200X_TRY(device_sm_count(&num_sms));
    const int tiles = ceil_div(p.M, 2 * GEMM_BM) * ceil_div(p.N, BN);
    const int pairs = std::min(tiles, num_sms / 2);
    gemm2_bf16_tn_kernel<BN><<<2 * pairs, GEMM_THREADS, L::TOTAL, stream>>>(tmA, tmB, tmC, tmCtail, p);
    B200X_CUDA_TRY(cudaGetLastError());
    return B200X_OK;
}

}  // namespace b200x

using namespace b200x;

extern "C" int b200x_gemm_bf16(const void* d_a, int lda, const void* d_w, int ldw, int M, int N, int K, int block_n,
                               void* d_out, int ldc, int out_mode, const float* d_bias, int act_gelu,
                               const float* d_resid, const float* d_pe, int group_in, int group_out, int group_off,
                               int reverse, void* stream) {
    B200X_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
    B200X_REQUIRE(N % 16 == 0, "gemm: N=%d must be a multiple of 16", N);
    B200X_REQUIRE(d_bias == nullptr || (reinterpret_cast<uintptr_t>(d_bias) & 15) == 0, "gemm: bias not 16-byte aligned");
    B200X_REQUIRE(lda % 8 == 0 && ldw % 8 == 0, "gemm: lda=%d / ldw=%d must be multiples of 8 (16-byte rows)", lda, ldw);
    B200X_REQUIRE(out_mode >= B200X_GEMM_OUT_BF16 && out_mode <= B200X_GEMM_OUT_F32_TOKEN, "gemm: bad out_mode %d", out_mode);
    B200X_REQUIRE(out_mode != B200X_GEMM_OUT_F32_RESID || (d_resid != nullptr && d_resid == d_out),
                  "gemm: the residual epilogue accumulates in place (TMA reduce-add): d_resid must equal d_out");
    B200X_REQUIRE(out_mode != B200X_GEMM_OUT_F32_RESID || act_gelu == 0, "gemm: GELU is not available with the residual epilogue");
    B200X_REQUIRE(ldc % (out_mode == B200X_GEMM_OUT_BF16 ? 8 : 4) == 0, "gemm: ldc=%d not 16-byte aligned", ldc);
    B200X_REQUIRE(out_mode != B200X_GEMM_OUT_F32_RESID || block_n % 32 == 0, "gemm: residual epilogue needs block_n %% 32 == 0");
    CUtensorMap tmA, tmB, tmC, tmCtail;
    const uint64_t da[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(M)};
    const uint64_t sa[1] = {static_cast<uint64_t>(lda) * 2};
    const uint32_t ba[2] = {GEMM_BK, GEMM_BM};
    B200X_TRY(make_tmap_bf16(&tmA, d_a, 2, da, sa, ba));
    const uint64_t dw[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(N)};
    const uint64_t sw[1] = {static_cast<uint64_t>(ldw) * 2};
    // CTA-pair kernel (256 x BN tiles, each CTA loads half of the B tile) for everything but the token epilogue
    const bool pair = out_mode != B200X_GEMM_OUT_F32_TOKEN && block_n != 128 && M > GEMM_BM;
    const uint32_t bw[2] = {GEMM_BK, static_cast<uint32_t>(pair ? block_n / 2 : block_n)};
    B200X_TRY(make_tmap_bf16(&tmB, d_w, 2, dw, sw, bw));
    // output maps: 32-row slabs, 128 bytes wide (64 bf16 / 32 fp32), plus a dense 16-column bf16 tail map
    const uint64_t dc[2] = {static_cast<uint64_t>(N), static_cast<uint64_t>(M)};
    if (out_mode == B200X_GEMM_OUT_F32_RESID) {
        const uint64_t sc[1] = {static_cast<uint64_t>(ldc) * 4};
        const uint32_t bc[2] = {32, 32};
        B200X_TRY(make_tmap(&tmC, d_out, 4, 2, dc, sc, bc, 1));
        tmCtail = tmC;
    } else if (out_mode == B200X_GEMM_OUT_BF16) {
        const uint64_t sc[1] = {static_cast<uint64_t>(ldc) * 2};
        const uint32_t bc[2] = {64, 32}, bt[2] = {16, 32};
        B200X_TRY(make_tmap(&tmC, d_out, 2, 2, dc, sc, bc, 1));
        B200X_TRY(make_tmap(&tmCtail, d_out, 2, 2, dc, sc, bt, 0));
    } else {
        tmC = tmA;            // unused in token mode
        tmCtail = tmA;
    }
    GemmParams p{M, N, K, d_out, ldc, out_mode, d_bias, act_gelu, d_pe, group_in, group_out, group_off, nullptr, (pair && reverse) ? 1 : 0};
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    // wide bf16 output of a narrow K (QKV, fc1): the A-stationary kernel, from four row tiles per CTA pair
    // upwards (its work unit is a whole row tile: below that the last, partly filled wave costs more than the saved operand
    // traffic gains); QKV 244 vs 267 us at 229 copies, 75.0 vs 78.2 us at 64; fc1 (epilogue-bound) 293 vs 297 us
    // (profiles/r02_p_kernel_bench.txt)
    int num_sms_d = 0;
    B200X_TRY(device_sm_count(&num_sms_d));
    const bool astat = pair && out_mode == B200X_GEMM_OUT_BF16 && (block_n == 192 || block_n == 208) && K % GEMM_BK == 0 &&
                       K <= GEMM_AS_MAX_KB * GEMM_BK && N >= 4 * block_n && ceil_div(M, 2 * GEMM_BM) >= 4 * (num_sms_d / 2);
    if (astat && block_n == 192) return launch_gemm2_astat<192, 8>(tmA, tmB, tmC, tmCtail, p, s);
    if (astat && block_n == 208) return act_gelu ? launch_gemm2_astat<208, B200X_FC1_EPI_WARPS>(tmA, tmB, tmC, tmCtail, p, s)
                                                 : launch_gemm2_astat<208, 8>(tmA, tmB, tmC, tmCtail, p, s);
    if (pair) {
        switch (block_n) {
            case 192: return launch_gemm2<192>(tmA, tmB, tmC, tmCtail, p, s);
            case 208: return launch_gemm2<208>(tmA, tmB, tmC, tmCtail, p, s);
            case 256: return launch_gemm2<256>(tmA, tmB, tmC, tmCtail, p, s);
            default: break;
        }
    }
    switch (block_n) {
        case 128: return launch_gemm<128>(tmA, tmB, tmC, tmCtail, p, s);
        case 192: return launch_gemm<192>(tmA, tmB, tmC, tmCtail, p, s);
        case 208: return launch_gemm<208>(tmA, tmB, tmC, tmCtail, p, s);
        case 256: return launch_gemm<256>(tmA, tmB, tmC, tmCtail, p, s);
        default: return set_error(B200X_ERR_INVALID, "gemm: unsupported block_n %d (128/192/208/256)", block_n);
    }
}

extern "C" int b200x_gemm_resid_ln_bf16(const void* d_a, int lda, const void* d_w, int ldw, int M, int N, int K, float* d_x, int ldx,
                                        const float* d_bias, const float* d_gamma, const float* d_beta, float eps, void* d_h, int ldh,
                                        int reverse, void* stream) {
    B200X_REQUIRE(d_a && d_w && d_x && d_gamma && d_beta && d_h, "gemm_resid_ln: NULL argument");
    B200X_REQUIRE(M > 0 && K > 0, "gemm_resid_ln: empty problem");
    B200X_REQUIRE(N % 128 == 0 && N <= 384, "gemm_resid_ln: N=%d (the LayerNorm width) must be a multiple of 128 up to 384", N);
    B200X_REQUIRE(K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0, "gemm_resid_ln: K / lda / ldw must be multiples of 8");
    B200X_REQUIRE(ldx % 4 == 0 && ldx >= N && ldh % 4 == 0 && ldh >= N, "gemm_resid_ln: ldx / ldh must be multiples of 4 and >= N");
    B200X_REQUIRE((reinterpret_cast<uintptr_t>(d_x) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_h) & 7) == 0 &&
                  (reinterpret_cast<uintptr_t>(d_gamma) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_beta) & 15) == 0 &&
                  (d_bias == nullptr || (reinterpret_cast<uintptr_t>(d_bias) & 15) == 0), "gemm_resid_ln: misaligned pointer");
    constexpr int BN = 192;
    CUtensorMap tmA, tmB, tmC;
    const uint64_t da[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(M)};
    const uint64_t sa[1] = {static_cast<uint64_t>(lda) * 2};
    const uint32_t ba[2] = {GEMM_BK, GEMM_BM};
    B200X_TRY(make_tmap_bf16(&tmA, d_a, 2, da, sa, ba));
    const uint64_t dw[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(N)};
    const uint64_t sw[1] = {static_cast<uint64_t>(ldw) * 2};
    const uint32_t bw[2] = {GEMM_BK, BN / 2};
    B200X_TRY(make_tmap_bf16(&tmB, d_w, 2, dw, sw, bw));
    const uint64_t dc[2] = {static_cast<uint64_t>(N), static_cast<uint64_t>(M)};
    const uint64_t sc[1] = {static_cast<uint64_t>(ldx) * 4};
    const uint32_t bc[2] = {32, 32};
    B200X_TRY(make_tmap(&tmC, d_x, 4, 2, dc, sc, bc, 1));
    GemmParams p{M, N, K, d_x, ldx, B200X_GEMM_OUT_F32_RESID, d_bias, 0, nullptr, 0, 0, 0, nullptr, reverse ? 1 : 0};
    GemmLnTail q{d_x, ldx, d_gamma, d_beta, eps, reinterpret_cast<__nv_bfloat16*>(d_h), ldh};
    // short K (attention projection): the load-add-store epilogue - its plain stores leave about half of the row tile in L2 for
    // the tail (303 us against 320 us with reduce-adds, 211 + 107 us as two kernels); long K (fc2): the reduce-add epilogue with
    // its five-stage ring (381 us against 425-449 us; 310 + 107 us as two kernels) - profiles/r02_h_gemm_resid_ln.txt
    if (K <= 512) return launch_gemm2_resid_ln<BN, true>(tmA, tmB, tmC, p, q, static_cast<cudaStream_t>(stream));
    return launch_gemm2_resid_ln<BN, false>(tmA, tmB, tmC, p, q, static_cast<cudaStream_t>(stream));
}

/* The A-stationary kernel on its own (b200x_gemm_bf16 picks it for wide bf16 outputs of a narrow K at large M): bf16 output,
 * optional bias / GELU, K a multiple of 64 up to 384, column tiles of 192 or 208. */
extern "C" int b200x_gemm_bf16_astationary(const void* d_a, int lda, const void* d_w, int ldw, int M, int N, int K, int block_n,
                                           void* d_out, int ldc, const float* d_bias, int act_gelu, int reverse, void* stream) {
    B200X_REQUIRE(block_n == 192 || block_n == 208, "gemm_astationary: block_n %d unsupported (192 / 208)", block_n);
    B200X_REQUIRE(d_a && d_w && d_out, "gemm_astationary: NULL argument");
    B200X_REQUIRE(M > 0 && N > 0 && N % 16 == 0, "gemm_astationary: bad problem M=%d N=%d", M, N);
    B200X_REQUIRE(K % GEMM_BK == 0 && K >= GEMM_BK && K <= GEMM_AS_MAX_KB * GEMM_BK, "gemm_astationary: K=%d must be a multiple of 64 up to 384", K);
    B200X_REQUIRE(lda % 8 == 0 && ldw % 8 == 0 && ldc % 8 == 0, "gemm_astationary: leading dimensions must be multiples of 8");
    B200X_REQUIRE(d_bias == nullptr || (reinterpret_cast<uintptr_t>(d_bias) & 15) == 0, "gemm_astationary: bias not 16-byte aligned");
    CUtensorMap tmA, tmB, tmC, tmCtail;
    const uint64_t da[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(M)};
    const uint64_t sa[1] = {static_cast<uint64_t>(lda) * 2};
    const uint32_t ba[2] = {GEMM_BK, GEMM_BM};
    B200X_TRY(make_tmap_bf16(&tmA, d_a, 2, da, sa, ba));
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
    if (block_n == 208) return act_gelu ? launch_gemm2_astat<208, B200X_FC1_EPI_WARPS>(tmA, tmB, tmC, tmCtail, p, static_cast<cudaStream_t>(stream))
                                        : launch_gemm2_astat<208, 8>(tmA, tmB, tmC, tmCtail, p, static_cast<cudaStream_t>(stream));
    return launch_gemm2_astat<192, 8>(tmA, tmB, tmC, tmCtail, p, static_cast<cudaStream_t>(stream));
}

/* Token-mode GEMM whose A operand is M-major: d_img bf16 [batch][K][128] (one 128-row tile per batch element, rows = the
 * contiguous index), out_row = b * group_out + group_off + m, out = act(A . W^T + bias) + pe[m].  The spectral tokenizer reads
 * the [time][mel] image the temporal tokenizer uses, so the transposed [mel][time] copy never has to be written. */
extern "C" int b200x_gemm_tokens_mmajor(const void* d_img, int batch, int K, const void* d_w, int ldw, int N, float* d_out, int ldc,
                                        const float* d_bias, int act_gelu, const float* d_pe, int group_out, int group_off,
                                        void* stream) {
    B200X_REQUIRE(d_img && d_w && d_out && d_pe, "gemm_tokens_mmajor: NULL argument");
    B200X_REQUIRE(batch > 0 && K > 0 && K % 8 == 0 && ldw % 8 == 0 && N > 0 && N % 4 == 0 && ldc % 4 == 0, "gemm_tokens_mmajor: bad sizes");
    B200X_REQUIRE(d_bias == nullptr || (reinterpret_cast<uintptr_t>(d_bias) & 15) == 0, "gemm_tokens_mmajor: bias not 16-byte aligned");
    constexpr int BN = 192;
    CUtensorMap tmA, tmB;
    const uint64_t da[3] = {GEMM_BM, static_cast<uint64_t>(K), static_cast<uint64_t>(batch)};
    const uint64_t sa[2] = {GEMM_BM * 2, static_cast<uint64_t>(K) * GEMM_BM * 2};
    const uint32_t ba[3] = {64, GEMM_BK, 1};
    B200X_TRY(make_tmap_bf16(&tmA, d_img, 3, da, sa, ba));
    const uint64_t dw[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(N)};
    const uint64_t sw[1] = {static_cast<uint64_t>(ldw) * 2};
    const uint32_t bw[2] = {GEMM_BK, BN};
    B200X_TRY(make_tmap_bf16(&tmB, d_w, 2, dw, sw, bw));
    GemmParams p{batch * GEMM_BM, N, K, d_out, ldc, B200X_GEMM_OUT_F32_TOKEN, d_bias, act_gelu, d_pe, GEMM_BM, group_out, group_off, nullptr, 0, 1};
    return launch_gemm<BN>(tmA, tmB, tmA, tmA, p, static_cast<cudaStream_t>(stream));
}
