// DSP stage of the perturbation hot path on sm_100a (n_fft = 2048, hop = 512, periodic Hann):
//   * stft_kernel ........... librosa-style centred STFT (zero padding) -> complex64 [frame][bin]       (once per track)
//   * istft_masked_kernel ... per perturbed copy: mask generated on the fly in the load stage (occlusion rectangle or
//                             per-bin band gain), inverse real FFT, synthesis window, overlap-add kept in REGISTERS while a
//                             warp walks a strip of consecutive frames, window-sum-square normalisation, coalesced stores
//   * mel_db_kernel ......... classifier front-end: reflect-padded STFT -> power -> HTK mel (sparse triangular filters)
//                             -> 10 log10 -> [copy][frame][mel] + per-CTA maxima (for the top_db clamp)
//   * mel_stats_kernel ...... clamp + sum / sum-of-squares partials (deterministic two-level reduction, fp64)
//   * mel_resize_kernel ..... normalise ((x-mean)/(std+eps)), bilinear resize along time, bf16, written in both operand
//                             layouts of the tokenizer GEMMs ([time][freq] and [freq][time])
// All FFTs are warp-level (fft.cuh); HBM/L2 traffic is coalesced float2 / float4.
#include <algorithm>
#include <cmath>
#include <map>
#include <mutex>
#include <vector>

#include "common.h"
#include "fft.cuh"
#include "ptx.cuh"

namespace b200x {

constexpr int NFFT = 2048;
constexpr int HOP = 512;
constexpr int NBIN = 1025;
constexpr int DSP_WARPS = 4;                    // warps per CTA for the FFT kernels
constexpr int DSP_THREADS = DSP_WARPS * 32;
constexpr int DSP_SMEM = (DSP_WARPS * FFT_TILE + FFT_TWIDDLE) * 8;
constexpr int ISTFT_ROW = 1026;                 // float2 elements copied per spectrogram row (1025 bins + 1: 16-byte multiple)
constexpr int MEL_MAX_SLOTS = 96;               // bins per lane of the filterbank walk (33 on average at 128 mels)
#ifdef B200X_MEL_NO_STAGE
constexpr int MEL_SMEM = DSP_SMEM + DSP_WARPS * 8;
#else
constexpr int MEL_SMEM = DSP_SMEM + DSP_WARPS * 2048 * 4 + DSP_WARPS * 8;   // + one staged 2048-sample frame and one mbarrier per warp
#endif
constexpr int ISTFT_SMEM = DSP_SMEM + DSP_WARPS * 2 * ISTFT_ROW * 8 + DSP_WARPS * 2 * 8;

__global__ void init_tables_kernel() {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 2048) {
        g_tw2048[i] = make_float2(static_cast<float>(cospi(i / 1024.0)), static_cast<float>(sinpi(i / 1024.0)));
        g_hann[i] = static_cast<float>(0.5 - 0.5 * cospi(i / 1024.0));
    }
    if (i < 512) {
        float acc = 0.f;
        for (int j = 0; j < 4; ++j) {
            const float w = static_cast<float>(0.5 - 0.5 * cospi((i + 512 * j) / 1024.0));
            acc += w * w;
        }
        g_wss512[i] = acc;
        g_rwss512[i] = static_cast<float>(1.0 / static_cast<double>(acc));
    }
}

// the __device__ tables live once per device (module instance): initialise them on first use of EACH device
static int ensure_tables(cudaStream_t stream) {
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);                  // a second host thread must not launch before the tables are filled
    bool first = false;
    B200X_TRY(device_first_use(reinterpret_cast<const void*>(init_tables_kernel), &first));
    if (!first) return B200X_OK;
    init_tables_kernel<<<8, 256, 0, stream>>>();
    B200X_CUDA_TRY(cudaGetLastError());
    B200X_CUDA_TRY(cudaStreamSynchronize(stream));
    return B200X_OK;
}

// real 2048-sample frame packed as z[m] = x[2m] + i x[2m+1]; after the 1024-point FFT, X[k] for k = lane + 32 r
// (and X[1024] on lane 0) is recovered from Z[k], Z[1024-k] staged in the warp tile.
template <bool TABLE>
__device__ __forceinline__ void rfft_unpack(const float2 (&v)[32], float2* tile, int lane, const LaneTrig& trig, float2 (&X)[32], float2& xnyq) {
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 32; ++r) tile[lane + 32 * r] = v[r];
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 32; ++r) {
        const int k = lane + 32 * r;
        const float2 zk = v[r];
        const float2 zp = tile[(1024 - k) & 1023];
        const float er = 0.5f * (zk.x + zp.x), ei = 0.5f * (zk.y - zp.y);
        const float orr = 0.5f * (zk.y + zp.y), oi = -0.5f * (zk.x - zp.x);
        const float2 w = TABLE ? __ldg(&g_tw2048[k]) : twiddle2048(trig, r);     // W = cos - i sin
        X[r] = make_float2(er + orr * w.x + oi * w.y, ei - orr * w.y + oi * w.x);
    }
    const float2 z0 = tile[0];
    xnyq = make_float2(z0.x - z0.y, 0.f);
    __syncwarp();
}

// ------------------------------------------------------------------------------------------------ STFT (librosa)
__global__ void __launch_bounds__(DSP_THREADS)
stft_kernel(const float* __restrict__ y_all, long long n_samples, int n_frames, int reflect, float2* __restrict__ S_all,
            int stride, long long y_copy_stride, long long s_copy_stride) {
    extern __shared__ __align__(16) float2 dsp_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* __restrict__ y = y_all + static_cast<long long>(blockIdx.y) * y_copy_stride;        // blockIdx.y = wave of a batch
    float2* __restrict__ S = S_all + static_cast<long long>(blockIdx.y) * s_copy_stride;
    float2* tile = dsp_smem + warp * FFT_TILE;
    float2* tw = dsp_smem + DSP_WARPS * FFT_TILE;
    fft_fill_twiddles(tw);
    const LaneTrig trig = lane_trig(lane);
    for (int t = blockIdx.x * DSP_WARPS + warp; t < n_frames; t += gridDim.x * DSP_WARPS) {
        const long long base = static_cast<long long>(t) * HOP - NFFT / 2;
        float2 v[32];
#pragma unroll
        for (int r = 0; r < 32; ++r) {
            const int m = lane + 32 * r;
            long long j0 = base + 2 * m, j1 = j0 + 1;
            float a, b;
            if (reflect) {
                if (j0 < 0) j0 = -j0;
                if (j1 < 0) j1 = -j1;
                if (j0 >= n_samples) j0 = 2 * (n_samples - 1) - j0;
                if (j1 >= n_samples) j1 = 2 * (n_samples - 1) - j1;
                a = y[j0]; b = y[j1];
            } else {
                a = (j0 >= 0 && j0 < n_samples) ? y[j0] : 0.f;
                b = (j1 >= 0 && j1 < n_samples) ? y[j1] : 0.f;
            }
            const float2 w = hann_pair(trig, r);
            v[r] = make_float2(a * w.x, b * w.y);
        }
        fft1024_warp<false>(v, tile, tw, lane);
        float2 X[32], xn;
        rfft_unpack<false>(v, tile, lane, trig, X, xn);
        float2* row = S + static_cast<long long>(t) * stride;
#pragma unroll
        for (int r = 0; r < 32; ++r) row[lane + 32 * r] = X[r];
        if (lane == 0) row[1024] = xn;
    }
}

// ------------------------------------------------------------------------------------------------ masked iSTFT
struct IstftParams {
    const float2* S;          // [n_frames][stride]
    int stride;
    int n_frames;
    long long out_len;        // samples written per copy = hop * (n_frames - 1)
    long long out_stride;     // distance between copies in y
    float* y;                 // [copies][out_stride]
    const int* windows;       // mode 1: [copies][4] = t0, t1, f0, f1
    float occlusion_value;
    const float* gains;       // mode 2: [copies][NBIN]
    int mode;                 // 0 none, 1 occlusion rectangle, 2 per-bin gain, 3 keep only the rectangle
    int hops_per_strip;
    double* sumsq;            // optional [copies]: sum of squares of the written samples (for RMS matching)
    int copies_per_track;     // copy c reads the spectrogram of track c / copies_per_track ...
    long long track_stride;   // ... which starts track_stride complex values after the previous track's
    const int* frame_range;   // optional [copies][2]: classifier frames [ma, mb) that differ from the unperturbed track;
                              // only the samples those frames read are synthesised (iSTFT linearity)
};

// MODE is a template parameter: the mask is evaluated per bin inside the fully unrolled load stage, and a run-time mode
// cost an ISETP / BRA pair per element there (ncu: a quarter of the kernel's samples sat in that stage)
#ifndef B200X_ISTFT_MIN_CTAS
#define B200X_ISTFT_MIN_CTAS 1
#endif
#ifndef B200X_MEL_MIN_CTAS
#define B200X_MEL_MIN_CTAS 4
#endif
template <int MODE>
__global__ void __launch_bounds__(DSP_THREADS, B200X_ISTFT_MIN_CTAS)
istft_masked_kernel(IstftParams p) {
    extern __shared__ __align__(16) float2 dsp_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float2* tile = dsp_smem + warp * FFT_TILE;
    float2* tw = dsp_smem + DSP_WARPS * FFT_TILE;
    fft_fill_twiddles(tw);
    const int copy = blockIdx.y;
    const int strip = blockIdx.x * DSP_WARPS + warp;
    // padded hops [hp_a, hp_b) of this strip; valid output hops are 2 .. n_frames (restricted to what frames [ma, mb) read)
    int h_lo = 2, h_hi = p.n_frames;
    if (p.frame_range != nullptr) {
        const int ma = p.frame_range[2 * copy], mb = p.frame_range[2 * copy + 1];
        if (mb <= ma) return;
        h_lo = max(2, ma);
        h_hi = min(p.n_frames, mb + 2);
    }
    const int hp_a = h_lo + strip * p.hops_per_strip;
    const int hp_b = min(hp_a + p.hops_per_strip, h_hi + 1);
    if (hp_a >= hp_b) return;
    int t0 = 0, t1 = 0, f0 = 0, f1 = 0;
    if (MODE == 1 || MODE == 3 || MODE == 4) {
        const int4 w = *reinterpret_cast<const int4*>(p.windows + 4 * copy);
        t0 = w.x; t1 = w.y; f0 = w.z; f1 = w.w;
    }
    const float* gain = MODE == 2 ? p.gains + static_cast<long long>(copy) * NBIN : nullptr;
    float* yout = p.y + static_cast<long long>(copy) * p.out_stride;
    const float2* spec = p.S + static_cast<long long>(copy / p.copies_per_track) * p.track_stride;

    const LaneTrig trig = lane_trig(lane);
    float2 a0[8], a1[8], a2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a0[i] = a1[i] = a2[i] = make_float2(0.f, 0.f);
    float sq = 0.f;

    // spectrogram rows are staged through shared memory with 1-D TMA bulk copies, double-buffered per warp: the copy of
    // frame t + 1 is in flight while frame t is transformed (the load stage used to stall on L2 latency, 61 % of samples)
    float2* rowbuf = dsp_smem + DSP_WARPS * FFT_TILE + FFT_TWIDDLE + warp * 2 * ISTFT_ROW;
    uint64_t* rbar = reinterpret_cast<uint64_t*>(dsp_smem + DSP_WARPS * FFT_TILE + FFT_TWIDDLE + DSP_WARPS * 2 * ISTFT_ROW) + warp * 2;
    const int t_first = max(hp_a - 3, 0);
    if (lane == 0) {
        mbar_init(&rbar[0], 1);
        mbar_init(&rbar[1], 1);
        fence_barrier_init();
        if (t_first < p.n_frames) {
            mbar_expect_tx(&rbar[0], ISTFT_ROW * 8);
            bulk_load_1d(rowbuf, spec + static_cast<long long>(t_first) * p.stride, ISTFT_ROW * 8, &rbar[0]);
        }
    }
    __syncwarp();

    for (int t = t_first; t < hp_b; ++t) {
        float2 v[32];
        const int it = t - t_first;
        if (lane == 0 && t + 1 < hp_b && t + 1 < p.n_frames) {      // the other buffer was read in the previous iteration
            uint64_t* nb = &rbar[(it + 1) & 1];
            mbar_expect_tx(nb, ISTFT_ROW * 8);
            bulk_load_1d(rowbuf + ((it + 1) & 1) * ISTFT_ROW, spec + static_cast<long long>(t + 1) * p.stride, ISTFT_ROW * 8, nb);
        }
        if (t < p.n_frames) {
            mbar_wait(&rbar[it & 1], (it >> 1) & 1);
            const float2* row = rowbuf + (it & 1) * ISTFT_ROW;
            const bool t_in = (t >= t0 && t < t1);
#pragma unroll
            for (int r = 0; r < 32; ++r) {
                const int k = lane + 32 * r, kp = 1024 - k;
                float2 xk = row[k];
                float2 xp = row[kp];
                if (MODE == 1) {
                    if (t_in && k >= f0 && k < f1) xk = make_float2(p.occlusion_value, 0.f);
                    if (t_in && kp >= f0 && kp < f1) xp = make_float2(p.occlusion_value, 0.f);
                } else if (MODE == 3) {
                    if (!(t_in && k >= f0 && k < f1)) xk = make_float2(0.f, 0.f);
                    if (!(t_in && kp >= f0 && kp < f1)) xp = make_float2(0.f, 0.f);
                } else if (MODE == 4) {                  // RISE: windows row = (seed, mask index, keep threshold, -)
                    const uint32_t key = rise_mask_key(static_cast<uint32_t>(t0), static_cast<uint32_t>(t1));
                    const uint32_t cell = static_cast<uint32_t>(t) * NBIN;
                    if (!rise_keep(key, cell + k, static_cast<uint32_t>(f0))) xk = make_float2(0.f, 0.f);
                    if (!rise_keep(key, cell + kp, static_cast<uint32_t>(f0))) xp = make_float2(0.f, 0.f);
                } else if (MODE == 2) {
                    const float gk = __ldg(&gain[k]), gp = __ldg(&gain[kp]);
                    xk.x *= gk; xk.y *= gk; xp.x *= gp; xp.y *= gp;
                }
                if (k == 0) xk.y = 0.f;                 // irfft ignores the imaginary part of DC ...
                if (kp == 1024) xp.y = 0.f;             // ... and of Nyquist
                const float er = 0.5f * (xk.x + xp.x), ei = 0.5f * (xk.y - xp.y);
                const float dr = 0.5f * (xk.x - xp.x), di = 0.5f * (xk.y + xp.y);
                const float2 w = twiddle2048(trig, r);   // e^{+i theta} = (cos, +sin)
                const float orr = dr * w.x - di * w.y, oi = dr * w.y + di * w.x;
                v[r] = make_float2(er - oi, ei + orr);
            }
            fft1024_warp<true>(v, tile, tw, lane);
#pragma unroll
            for (int r = 0; r < 32; ++r) {
                const float2 w = hann_pair(trig, r);
                v[r] = make_float2(v[r].x * w.x * (1.0f / 1024.0f), v[r].y * w.y * (1.0f / 1024.0f));
            }
        } else {
#pragma unroll
            for (int r = 0; r < 32; ++r) v[r] = make_float2(0.f, 0.f);
        }
        // overlap-add: frame chunk q = r / 8 lands on padded hop t + q; hop t is now complete
        if (t >= hp_a) {
            const bool steady = (t >= 3 && t <= p.n_frames - 1);
            const long long n_base = static_cast<long long>(t - 2) * HOP;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int slot = 2 * lane + 64 * i;
                float2 o = make_float2(a0[i].x + v[i].x, a0[i].y + v[i].y);
                if (steady) {
                    // steady state (all four overlapping frames exist): the envelope is the 512-periodic table; multiply by
                    // its reciprocal (rounded once from double) instead of dividing
                    const float2 rw = *reinterpret_cast<const float2*>(&g_rwss512[slot]);
                    o.x *= rw.x; o.y *= rw.y;
                } else {
                    float2 wss = make_float2(0.f, 0.f);
                    for (int j = 0; j < 4; ++j) {
                        const int tf = t - j;
                        if (tf >= 0 && tf < p.n_frames) {
                            const float2 w = *reinterpret_cast<const float2*>(&g_hann[j * HOP + slot]);
                            wss.x += w.x * w.x; wss.y += w.y * w.y;
                        }
                    }
                    if (wss.x > 1.17549435e-38f) o.x /= wss.x;
                    if (wss.y > 1.17549435e-38f) o.y /= wss.y;
                }
                if (n_base + slot < p.out_len) {
                    *reinterpret_cast<float2*>(yout + n_base + slot) = o;
                    sq += o.x * o.x + o.y * o.y;
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            a0[i] = make_float2(a1[i].x + v[8 + i].x, a1[i].y + v[8 + i].y);
            a1[i] = make_float2(a2[i].x + v[16 + i].x, a2[i].y + v[16 + i].y);
            a2[i] = v[24 + i];
        }
        __syncwarp();
    }
    if (p.sumsq != nullptr) {
        double d = static_cast<double>(sq);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
        if (lane == 0) atomicAdd(p.sumsq + copy, d);
    }
}

// ------------------------------------------------------------------------------------------------ mel front-end
struct MelParams {
    const float* y;            // [copies][y_stride]
    long long y_stride;
    long long n_samples;
    const double* sumsq;       // optional RMS matching: per-copy sum of squares of y, and the reference RMS
    double ref_rms;
    const double* ref_rms_arr; // optional [copies]: per-copy reference RMS (several tracks in one launch); < 0 = leave the copy unscaled
    long long rms_count;
    int n_frames;
    int n_mels;                // <= 128, multiple of 32
    // triangular filterbank by SEGMENTS between consecutive centre frequencies: a bin of segment j feeds filter j with its
    // rising weight (.x) and filter j - 1 with its falling weight (.y), so mel[m] = U[m] + D[m + 1] with two sums per segment.
    // Every lane owns a CONTIGUOUS run of whole (non-empty) segments, balanced by width; its bins are "slots" 0 .. mel_slots - 1.
    const float2* lane_weights; // [mel_slots][32] (rising, falling) weight of slot i of lane l (slot-major: one coalesced 256-byte
                                // load per slot); the sign bit of .x marks the LAST bin of a segment; (+0, 0) past the lane's run
    const int2* lane_info;      // [32] (first bin, first compacted segment id) of each lane
    const int2* filter_segs;    // [n_mels] compacted ids of the segments whose U / D sums make up filter m (-1 = empty segment)
    int mel_slots;              // multiple of 8
    float amin;
    float* db;                 // [copies][db_frames][n_mels]; row (t - ma) when frame_range is given
    float* cta_max;            // [copies][gridDim.x]
    int frames_per_cta;
    int db_frames;             // frame capacity per copy of db
    const int* frame_range;    // optional [copies][2] = [ma, mb): only these frames are computed
};

__global__ void __launch_bounds__(DSP_THREADS, B200X_MEL_MIN_CTAS)      // 4 -> 128 registers (no spills); 74 KB of shared memory still means three CTAs per SM,
                                                                        // but the tighter allocation measured 2-3 % faster (profiles/r02_k_dsp_variants.txt)
mel_db_kernel(MelParams p) {
    extern __shared__ __align__(16) float2 dsp_smem[];
    __shared__ float s_max[DSP_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float2* tile = dsp_smem + warp * FFT_TILE;
    float2* tw = dsp_smem + DSP_WARPS * FFT_TILE;
    const int copy = blockIdx.y;
    {   // CTAs past this copy's frame range have nothing to do (the grid is sized for the widest range of the launch)
        int lo = 0, hi = p.n_frames;
        if (p.frame_range != nullptr) { lo = p.frame_range[2 * copy]; hi = p.frame_range[2 * copy + 1]; }
        if (lo + static_cast<int>(blockIdx.x) * p.frames_per_cta >= hi) {
            if (threadIdx.x == 0) p.cta_max[static_cast<long long>(copy) * gridDim.x + blockIdx.x] = -INFINITY;
            return;
        }
    }
    fft_fill_twiddles(tw);
    float* pw = reinterpret_cast<float*>(tile);
    const float* y = p.y + static_cast<long long>(copy) * p.y_stride;
    float gain = 1.0f;
    if (p.sumsq != nullptr) {                         // match_rms (src/dsp_band_ops.py:228-233), float64 like the reference
        const double r_x = sqrt(p.sumsq[copy] / static_cast<double>(p.rms_count) + 1e-8);
        const double ref = p.ref_rms_arr != nullptr ? p.ref_rms_arr[copy] : p.ref_rms;
        if (!(r_x < 1e-8) && ref >= 0.0) gain = static_cast<float>(ref / r_x);
    }
    const LaneTrig trig{};      // unused: here the table loads of the window and the unpack twiddles measured faster
                                // (662 us vs 782 us per 64 sparse copies) than computing them
    float vmax = -INFINITY;
    const int2 lane_run = __ldg(&p.lane_info[lane]);
    const int lane_bin0 = lane_run.x, lane_seg0 = lane_run.y;
    int f_lo = 0, f_hi = p.n_frames;
    if (p.frame_range != nullptr) { f_lo = p.frame_range[2 * copy]; f_hi = p.frame_range[2 * copy + 1]; }
    const int f_begin = f_lo + blockIdx.x * p.frames_per_cta;
    const int f_end = min(f_begin + p.frames_per_cta, f_hi);
    // Interior frames (2048 contiguous, 16-byte aligned samples) are staged through shared memory with a 1-D TMA bulk copy
    // per warp: the copy of the warp's NEXT frame is issued as soon as the current one sits in registers and lands during
    // the FFT, so the load stage no longer waits on L2 latency.  Edge frames (reflect padding) take the direct path.
    float* fbuf = reinterpret_cast<float*>(dsp_smem + DSP_WARPS * FFT_TILE + FFT_TWIDDLE) + warp * NFFT;
#ifdef B200X_MEL_NO_STAGE
    uint64_t* fbar = reinterpret_cast<uint64_t*>(dsp_smem + DSP_WARPS * FFT_TILE + FFT_TWIDDLE) + warp;
#else
    uint64_t* fbar = reinterpret_cast<uint64_t*>(reinterpret_cast<float*>(dsp_smem + DSP_WARPS * FFT_TILE + FFT_TWIDDLE) + DSP_WARPS * NFFT) + warp;
#endif
#ifdef B200X_MEL_NO_STAGE
    const bool can_stage = false;
#else
    const bool can_stage = ((reinterpret_cast<uintptr_t>(y) & 15) == 0);
#endif
    auto interior_of = [&](int t) {
        const long long b = static_cast<long long>(t) * HOP - NFFT / 2;
        return b >= 0 && b + NFFT <= p.n_samples;
    };
    if (lane == 0) {
        mbar_init(fbar, 1);
        fence_barrier_init();
    }
    __syncwarp();
    bool pending = false;
    uint32_t staged = 0;
    {
        const int t0 = f_begin + warp;
        if (t0 < f_end && can_stage && interior_of(t0)) {
            pending = true;
            if (lane == 0) {
                mbar_expect_tx(fbar, NFFT * 4);
                bulk_load_1d(fbuf, y + static_cast<long long>(t0) * HOP - NFFT / 2, NFFT * 4, fbar);
            }
        }
    }
    for (int t = f_begin + warp; t < f_end; t += DSP_WARPS) {
        const long long base = static_cast<long long>(t) * HOP - NFFT / 2;
        float2 v[32];
        const bool interior = (base >= 0) && (base + NFFT <= p.n_samples);
        if (pending) {
            mbar_wait(fbar, staged & 1);
            ++staged;
#pragma unroll
            for (int r = 0; r < 32; ++r) {
                const int m = lane + 32 * r;
                const float2 s = *reinterpret_cast<const float2*>(fbuf + 2 * m);
                const float2 w = *reinterpret_cast<const float2*>(&g_hann[2 * m]);
                v[r] = make_float2(s.x * gain * w.x, s.y * gain * w.y);
            }
            __syncwarp();                              // every lane has read the buffer: it may be refilled
        } else {
#pragma unroll
        for (int r = 0; r < 32; ++r) {
            const int m = lane + 32 * r;
            float2 s;
            if (interior) {
                s = *reinterpret_cast<const float2*>(y + base + 2 * m);
            } else {                                   // reflect padding (torch.stft center=True, pad_mode='reflect')
                long long j0 = base + 2 * m, j1 = j0 + 1;
                if (j0 < 0) j0 = -j0;
                if (j1 < 0) j1 = -j1;
                if (j0 >= p.n_samples) j0 = 2 * (p.n_samples - 1) - j0;
                if (j1 >= p.n_samples) j1 = 2 * (p.n_samples - 1) - j1;
                s = make_float2(y[j0], y[j1]);
            }
            const float2 w = *reinterpret_cast<const float2*>(&g_hann[2 * m]);
            v[r] = make_float2(s.x * gain * w.x, s.y * gain * w.y);
        }
        }
        {
            const int tn = t + DSP_WARPS;
            pending = tn < f_end && can_stage && interior_of(tn);
            if (pending && lane == 0) {
                mbar_expect_tx(fbar, NFFT * 4);
                bulk_load_1d(fbuf, y + static_cast<long long>(tn) * HOP - NFFT / 2, NFFT * 4, fbar);
            }
        }
        fft1024_warp<false>(v, tile, tw, lane);
        float2 X[32], xn;
        rfft_unpack<true>(v, tile, lane, trig, X, xn);
#pragma unroll
        for (int r = 0; r < 32; ++r) pw[lane + 32 * r] = X[r].x * X[r].x + X[r].y * X[r].y;
        if (lane == 0) pw[1024] = xn.x * xn.x;
        __syncwarp();
        float* out = p.db + (static_cast<long long>(copy) * p.db_frames + (t - f_lo)) * p.n_mels;
        // filterbank: every lane walks its contiguous run of bins once.  The weights of a slot come from one coalesced load,
        // all loads of a batch of eight slots are independent of the sums, and a segment's (U, D) pair is flushed when the
        // sign-bit flag says its last bin has been added - the per-segment loops this replaces (variable trip counts, one
        // lane-scattered weight load per step) held 47 % of the kernel's stall samples on their dependent loads.
        // Same additions in the same order as before: results are bit-identical.
        float* segU = pw + 1056;
        float* segD = segU + 160;
        {
            int j = lane_seg0;
            float au = 0.f, ad = 0.f;
            for (int base = 0; base < p.mel_slots; base += 8) {
                float2 w[8];
                float x[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    w[u] = __ldg(&p.lane_weights[(base + u) * 32 + lane]);
                    x[u] = pw[min(lane_bin0 + base + u, NBIN - 1)];
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    au = fmaf(fabsf(w[u].x), x[u], au);
                    ad = fmaf(w[u].y, x[u], ad);
                    if (__float_as_int(w[u].x) < 0) {          // last bin of segment j
                        segU[j] = au;
                        segD[j] = ad;
                        au = 0.f;
                        ad = 0.f;
                        ++j;
                    }
                }
            }
        }
        __syncwarp();
        for (int f = lane; f < p.n_mels; f += 32) {
            const int2 m = __ldg(&p.filter_segs[f]);
            const float acc = (m.x >= 0 ? segU[m.x] : 0.f) + (m.y >= 0 ? segD[m.y] : 0.f);
            const float d = 10.0f * log10f(fmaxf(acc, p.amin));
            out[f] = d;
            vmax = fmaxf(vmax, d);
        }
        __syncwarp();
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    if (lane == 0) s_max[warp] = vmax;
    __syncthreads();
    if (threadIdx.x == 0) {
        float m = s_max[0];
        for (int i = 1; i < DSP_WARPS; ++i) m = fmaxf(m, s_max[i]);
        p.cta_max[static_cast<long long>(copy) * gridDim.x + blockIdx.x] = m;
    }
}

// Where a copy's dB rows live: frames [ma, mb) in its own compact buffer, every other frame in the track's baseline
// (those frames are bit-identical to the unperturbed track, so they are computed once per track).
struct DbView {
    const float* own;          // [db_frames][n_mels] of this copy
    const float* base;         // [n_frames][n_mels] baseline or nullptr
    int ma, mb, n_mels;
    __device__ __forceinline__ const float* row(int t) const {
        return (base == nullptr || (t >= ma && t < mb)) ? own + static_cast<long long>(t - ma) * n_mels
                                                        : base + static_cast<long long>(t) * n_mels;
    }
};

struct StatsParams {
    const float* db;           // [copies][db_frames][n_mels]
    int db_frames, n_frames, n_mels;
    const float* cta_max;
    int n_cta_max;
    float top_db;
    const float* base;         // optional baseline dB [n_frames][n_mels]
    const float* base_premax;  // [n_frames + 1]: max over baseline frames < m
    const float* base_sufmax;  // [n_frames + 1]: max over baseline frames >= m
    const int* frame_range;    // [copies][2] (required when base != nullptr)
    double2* partial;
    float* floor_out;
};

// clamp at (max - top_db) and reduce sum / sum of squares: partial[copy][block] = (sum, sumsq), fp64
__global__ void __launch_bounds__(256)
mel_stats_kernel(StatsParams p) {
    __shared__ float s_floor;
    __shared__ double s_a[8], s_b[8];
    const int copy = blockIdx.y;
    DbView v;
    v.own = p.db + static_cast<long long>(copy) * p.db_frames * p.n_mels;
    v.base = p.base; v.n_mels = p.n_mels; v.ma = 0; v.mb = p.n_frames;
    if (p.base != nullptr) { v.ma = p.frame_range[2 * copy]; v.mb = p.frame_range[2 * copy + 1]; }
    if (threadIdx.x == 0) {
        float m = -INFINITY;
        for (int i = 0; i < p.n_cta_max; ++i) m = fmaxf(m, p.cta_max[static_cast<long long>(copy) * p.n_cta_max + i]);
        if (p.base != nullptr) m = fmaxf(m, fmaxf(p.base_premax[v.ma], p.base_sufmax[v.mb]));
        s_floor = m - p.top_db;
        if (blockIdx.x == 0) p.floor_out[copy] = s_floor;
    }
    __syncthreads();
    const float fl = s_floor;
    const int per_row4 = p.n_mels / 4;
    const long long n4 = static_cast<long long>(p.n_frames) * per_row4;
    double a = 0.0, b = 0.0;
    for (long long i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int t = static_cast<int>(i / per_row4), c4 = static_cast<int>(i % per_row4);
        const float4 x = reinterpret_cast<const float4*>(v.row(t))[c4];
        const float v0 = fmaxf(x.x, fl), v1 = fmaxf(x.y, fl), v2 = fmaxf(x.z, fl), v3 = fmaxf(x.w, fl);
        a += static_cast<double>(v0) + v1 + v2 + v3;
        b += static_cast<double>(v0) * v0 + static_cast<double>(v1) * v1 + static_cast<double>(v2) * v2 +
             static_cast<double>(v3) * v3;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if ((threadIdx.x & 31) == 0) { s_a[threadIdx.x >> 5] = a; s_b[threadIdx.x >> 5] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double sa = 0.0, sb = 0.0;
        for (int i = 0; i < 8; ++i) { sa += s_a[i]; sb += s_b[i]; }
        p.partial[static_cast<long long>(copy) * gridDim.x + blockIdx.x] = make_double2(sa, sb);
    }
}

// normalise + bilinear resize along time (F.interpolate(mode='bilinear', align_corners=False); the mel axis keeps its
// size so its weights are exactly (1, 0)) -> bf16 in [time][mel] (temporal tokenizer operand) and [mel][time] (spectral)
struct ResizeParams {
    const float* db;           // [copies][db_frames][n_mels]
    int db_frames;
    const float* base;         // optional baseline dB
    const int* frame_range;
    const double2* partial;
    const float* floor_val;
    int n_partial;
    int n_frames, n_mels, out_t;
    int unbiased;
    float eps;
    __nv_bfloat16* img_t;      // [copies][out_t][n_mels]
    __nv_bfloat16* img_f;      // [copies][n_mels][ld_f]
    int ld_f;
};

__global__ void __launch_bounds__(256)
mel_resize_kernel(ResizeParams p) {
    __shared__ float s_mean, s_inv;
    __shared__ __align__(8) __nv_bfloat16 s_tile[64][128 + 4];
    const int copy = blockIdx.y;
    if (threadIdx.x == 0) {
        double sa = 0.0, sb = 0.0;
        for (int i = 0; i < p.n_partial; ++i) {
            const double2 v = p.partial[static_cast<long long>(copy) * p.n_partial + i];
            sa += v.x; sb += v.y;
        }
        const double n = static_cast<double>(p.n_frames) * p.n_mels;
        const double mean = sa / n;
        const double var = fmax((sb - n * mean * mean) / (p.unbiased ? n - 1.0 : n), 0.0);
        s_mean = static_cast<float>(mean);
        s_inv = 1.0f / (static_cast<float>(sqrt(var)) + p.eps);
    }
    __syncthreads();
    const float mean = s_mean, inv = s_inv, fl = p.floor_val[copy];
    const float scale = static_cast<float>(p.n_frames) / static_cast<float>(p.out_t);
    DbView v;
    v.own = p.db + static_cast<long long>(copy) * p.db_frames * p.n_mels;
    v.base = p.base; v.n_mels = p.n_mels; v.ma = 0; v.mb = p.n_frames;
    if (p.base != nullptr) { v.ma = p.frame_range[2 * copy]; v.mb = p.frame_range[2 * copy + 1]; }
    const int j0 = blockIdx.x * 64;
    // one thread = four consecutive mel bins of one output row: the source rows, the interpolation weights and the
    // own / baseline row selection are computed once per row and thread instead of once per element (the element-wise
    // form spent 62 % of its issue slots on integer index arithmetic), loads are 16 bytes, stores 8 bytes
    const int quads = p.n_mels >> 2;                     // n_mels is a multiple of 32
    for (int idx = threadIdx.x; idx < 64 * quads; idx += blockDim.x) {
        const int jj = idx / quads, f = (idx % quads) << 2;
        const int j = j0 + jj;
        if (j >= p.out_t) continue;
        float src = scale * (static_cast<float>(j) + 0.5f) - 0.5f;
        if (src < 0.f) src = 0.f;
        const int i0 = static_cast<int>(src);
        const int i1 = i0 + (i0 < p.n_frames - 1 ? 1 : 0);
        const float lam1 = src - static_cast<float>(i0), lam0 = 1.0f - lam1;
        const float4 a = *reinterpret_cast<const float4*>(v.row(i0) + f);
        const float4 b = *reinterpret_cast<const float4*>(v.row(i1) + f);
        const float o0 = lam0 * ((fmaxf(a.x, fl) - mean) * inv) + lam1 * ((fmaxf(b.x, fl) - mean) * inv);
        const float o1 = lam0 * ((fmaxf(a.y, fl) - mean) * inv) + lam1 * ((fmaxf(b.y, fl) - mean) * inv);
        const float o2 = lam0 * ((fmaxf(a.z, fl) - mean) * inv) + lam1 * ((fmaxf(b.z, fl) - mean) * inv);
        const float o3 = lam0 * ((fmaxf(a.w, fl) - mean) * inv) + lam1 * ((fmaxf(b.w, fl) - mean) * inv);
        const uint2 w = make_uint2(pack_bf16(o0, o1), pack_bf16(o2, o3));
        *reinterpret_cast<uint2*>(p.img_t + (static_cast<long long>(copy) * p.out_t + j) * p.n_mels + f) = w;
        if (p.img_f != nullptr) *reinterpret_cast<uint2*>(&s_tile[jj][f]) = w;
    }
    if (p.img_f == nullptr) return;                      // the consumer reads the [time][mel] image through an M-major descriptor
    __syncthreads();
    // transposed copy [mel][time]: one thread = four consecutive output frames of one mel bin (8-byte stores when aligned)
    const bool vec_ok = (p.ld_f % 4 == 0) && (j0 + 64 <= p.out_t);
    for (int idx = threadIdx.x; idx < 16 * p.n_mels; idx += blockDim.x) {
        const int f = idx >> 4, jq = (idx & 15) << 2;
        __nv_bfloat16* dst = p.img_f + (static_cast<long long>(copy) * p.n_mels + f) * p.ld_f + j0 + jq;
        if (vec_ok) {
            const uint32_t lo = static_cast<uint32_t>(__bfloat16_as_ushort(s_tile[jq][f])) | (static_cast<uint32_t>(__bfloat16_as_ushort(s_tile[jq + 1][f])) << 16);
            const uint32_t hi = static_cast<uint32_t>(__bfloat16_as_ushort(s_tile[jq + 2][f])) | (static_cast<uint32_t>(__bfloat16_as_ushort(s_tile[jq + 3][f])) << 16);
            *reinterpret_cast<uint2*>(dst) = make_uint2(lo, hi);
        } else {
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (j0 + jq + u < p.out_t) dst[u] = s_tile[jq + u][f];
        }
    }
}

// classifier frames touched by an occlusion window: a patch on STFT frames [t0, t1) changes samples
// [t0*hop - n_fft/2, (t1-1)*hop + n_fft/2), i.e. classifier (hop 512, n_fft 2048) frames [t0 - 3, t1 + 3); one more
// frame on each side covers the sample that the first / last frame reaches through reflect padding
__global__ void frame_ranges_kernel(const int* __restrict__ windows, int n, int n_frames, int* __restrict__ ranges) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int t0 = windows[4 * i], t1 = windows[4 * i + 1], f0 = windows[4 * i + 2], f1 = windows[4 * i + 3];
    int ma = 0, mb = 0;
    if (t1 > t0 && f1 > f0) { ma = max(t0 - 4, 0); mb = min(t1 + 4, n_frames); }
    ranges[2 * i] = ma;
    ranges[2 * i + 1] = mb;
}

// pre[m] = max over baseline frames < m, suf[m] = max over frames >= m  (m = 0 .. n_frames); one block
__global__ void base_maxima_kernel(const float* __restrict__ db, int n_frames, int n_mels, float* __restrict__ pre,
                                   float* __restrict__ suf) {
    extern __shared__ float s_fm[];                 // per-frame maxima
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
    for (int t = warp; t < n_frames; t += n_warps) {      // one warp per frame: coalesced row reads, shuffle reduction
        float m = -INFINITY;
        for (int f = lane; f < n_mels; f += 32) m = fmaxf(m, db[static_cast<long long>(t) * n_mels + f]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 0) s_fm[t] = m;
    }
    __syncthreads();
    // running maxima by two warps (prefix / suffix), each lane scanning one contiguous segment after a shuffle scan of
    // the segment maxima (max is exactly associative, so the result does not depend on the split)
    if (warp < 2) {
        const int seg = (n_frames + 31) / 32;
        const int a = min(lane * seg, n_frames), b = min(a + seg, n_frames);
        float m = -INFINITY;
        for (int t = a; t < b; ++t) m = fmaxf(m, s_fm[t]);
        if (warp == 0) {
            float carry = m;                               // inclusive scan of the segment maxima, then shift by one lane
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const float v = __shfl_up_sync(0xffffffffu, carry, o);
                if (lane >= o) carry = fmaxf(carry, v);
            }
            float run = __shfl_up_sync(0xffffffffu, carry, 1);
            if (lane == 0) run = -INFINITY;
            for (int t = a; t < b; ++t) { pre[t] = run; run = fmaxf(run, s_fm[t]); }
            if (lane == 31) pre[n_frames] = carry;         // maximum over every frame
        } else {
            float carry = m;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const float v = __shfl_down_sync(0xffffffffu, carry, o);
                if (lane + o < 32) carry = fmaxf(carry, v);
            }
            float run = __shfl_down_sync(0xffffffffu, carry, 1);
            if (lane == 31) run = -INFINITY;
            for (int t = b - 1; t >= a; --t) { run = fmaxf(run, s_fm[t]); suf[t] = run; }
            if (lane == 0) suf[n_frames] = -INFINITY;
        }
    }
}

// ref[i] = sqrt(mean(wave_i^2) + 1e-8) in float64 (match_rms reference level, src/dsp_band_ops.py:228-233), one CTA per wave;
// out[i * repeat .. (i + 1) * repeat) all receive it (one entry per perturbed copy of that track)
__global__ void __launch_bounds__(1024)
wave_rms_kernel(const float* __restrict__ waves, long long n_samples, long long stride, int repeat, double* __restrict__ out) {
    __shared__ double s_part[32];
    const float* w = waves + static_cast<long long>(blockIdx.x) * stride;
    double acc = 0.0;
    for (long long i = threadIdx.x; i < n_samples; i += blockDim.x) { const double v = w[i]; acc += v * v; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < static_cast<int>(blockDim.x >> 5); ++i) t += s_part[i];
        const double r = sqrt(t / static_cast<double>(n_samples) + 1e-8);
        for (int k = 0; k < repeat; ++k) out[static_cast<long long>(blockIdx.x) * repeat + k] = r;
    }
}

// y[b] = sum_i masks[b][i] * stems[i]  (LIME stem recombination, src/lime_explainer.py:283-301)
__global__ void mix_stems_kernel(const float* __restrict__ stems, long long n_samples, int n_stems,
                                 const unsigned char* __restrict__ masks, float* __restrict__ y, long long y_stride) {
    const int copy = blockIdx.y;
    for (long long i = blockIdx.x * blockDim.x + threadIdx.x; i < n_samples; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        float acc = 0.f;
        for (int s = 0; s < n_stems; ++s)
            if (masks[copy * n_stems + s]) acc += stems[s * n_samples + i];
        y[copy * y_stride + i] = acc;
    }
}

}  // namespace b200x

using namespace b200x;

extern "C" int b200x_stft(const float* d_wave, int64_t n_samples, int n_fft, int hop, int reflect_pad, void* d_spec,
                          int spec_stride, void* stream) {
    B200X_REQUIRE(n_fft == NFFT && hop == HOP, "stft: only n_fft=2048, hop=512 are built (got %d/%d)", n_fft, hop);
    B200X_REQUIRE(n_samples > NFFT / 2 && spec_stride >= NBIN, "stft: bad sizes");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    B200X_TRY(ensure_tables(s));
    B200X_TRY(ensure_kernel_smem(reinterpret_cast<const void*>(stft_kernel), DSP_SMEM));
    const int n_frames = 1 + static_cast<int>(n_samples / HOP);
    const int grid = std::min(ceil_div(n_frames, DSP_WARPS), 148 * 8);
    stft_kernel<<<grid, DSP_THREADS, DSP_SMEM, s>>>(d_wave, n_samples, n_frames, reflect_pad, reinterpret_cast<float2*>(d_spec), spec_stride, 0, 0);
    B200X_CUDA_TRY(cudaGetLastError());
    return B200X_OK;
}

// librosa.stft of `copies` equal-length waves in one launch (Griffin-Lim: rebuilt = stft(istft(.)) for a chunk of copies):
// wave c at d_waves + c * wave_stride floats, spectrum c at d_spec + c * spec_copy_stride complex values
extern "C" int b200x_stft_batch(const float* d_waves, int64_t n_samples, int64_t wave_stride, int copies, void* d_spec,
                                int spec_stride, int64_t spec_copy_stride, void* stream) {
    B200X_REQUIRE(d_waves && d_spec && copies > 0, "stft_batch: bad argument");
    B200X_REQUIRE(n_samples > NFFT / 2 && spec_stride >= NBIN && wave_stride >= n_samples, "stft_batch: bad sizes");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    B200X_TRY(ensure_tables(s));
    B200X_TRY(ensure_kernel_smem(reinterpret_cast<const void*>(stft_kernel), DSP_SMEM));
    const int n_frames = 1 + static_cast<int>(n_samples / HOP);
    B200X_REQUIRE(spec_copy_stride >= static_cast<int64_t>(n_frames) * spec_stride, "stft_batch: spec_copy_stride too small");
    dim3 grid(std::min(ceil_div(n_frames, DSP_WARPS), 148 * 8), copies);
    stft_kernel<<<grid, DSP_THREADS, DSP_SMEM, s>>>(d_waves, n_samples, n_frames, 0, reinterpret_cast<float2*>(d_spec), spec_stride, wave_stride,
                                                    spec_copy_stride);
    B200X_CUDA_TRY(cudaGetLastError());
    return B200X_OK;
}

extern "C" int b200x_istft_masked(const void* d_spec, int spec_stride, int n_frames, int copies, int mode,
                                  const int32_t* d_windows, float occlusion_value, const float* d_gains, float* d_y,
                                  int64_t y_stride, double* d_sumsq, const int32_t* d_frame_range, int max_range_frames,
                                  void* stream) {
    return b200x_istft_masked_tracks(d_spec, spec_stride, n_frames, copies, copies > 0 ? copies : 1, 0, mode, d_windows, occlusion_value,
                                     d_gains, d_y, y_stride, d_sumsq, d_frame_range, max_range_frames, stream);
}

extern "C" int b200x_istft_masked_tracks(const void* d_spec, int spec_stride, int n_frames, int copies, int copies_per_track,
                                         int64_t track_stride, int mode, const int32_t* d_windows, float occlusion_value,
                                         const float* d_gains, float* d_y, int64_t y_stride, double* d_sumsq,
                                         const int32_t* d_frame_range, int max_range_frames, void* stream) {
    B200X_REQUIRE(copies_per_track > 0 && track_stride >= 0 && (track_stride * 8) % 16 == 0, "istft: bad track layout");
    B200X_REQUIRE(mode >= 0 && mode <= 4, "istft: bad mode %d", mode);
    B200X_REQUIRE((mode != 1 && mode != 3 && mode != 4) || d_windows != nullptr, "istft: windows missing");
    B200X_REQUIRE(mode != 2 || d_gains != nullptr, "istft: gains missing");
    B200X_REQUIRE(n_frames >= 2 && copies > 0, "istft: bad sizes");
    B200X_REQUIRE(spec_stride >= ISTFT_ROW && spec_stride % 2 == 0 && (reinterpret_cast<uintptr_t>(d_spec) & 15) == 0,
                  "istft: spectrogram rows must be 16-byte aligned with stride >= %d complex values (got %d)", ISTFT_ROW, spec_stride);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    B200X_TRY(ensure_tables(s));
    B200X_TRY(ensure_kernel_smem(reinterpret_cast<const void*>(istft_masked_kernel<0>), ISTFT_SMEM));
    B200X_TRY(ensure_kernel_smem(reinterpret_cast<const void*>(istft_masked_kernel<1>), ISTFT_SMEM));
    B200X_TRY(ensure_kernel_smem(reinterpret_cast<const void*>(istft_masked_kernel<2>), ISTFT_SMEM));
    B200X_TRY(ensure_kernel_smem(reinterpret_cast<const void*>(istft_masked_kernel<3>), ISTFT_SMEM));
    B200X_TRY(ensure_kernel_smem(reinterpret_cast<const void*>(istft_masked_kernel<4>), ISTFT_SMEM));
    IstftParams p;
    p.S = reinterpret_cast<const float2*>(d_spec); p.stride = spec_stride; p.n_frames = n_frames;
    p.out_len = static_cast<long long>(HOP) * (n_frames - 1); p.out_stride = y_stride; p.y = d_y;
    p.windows = d_windows; p.occlusion_value = occlusion_value; p.gains = d_gains; p.mode = mode;
    p.sumsq = d_sumsq; p.frame_range = d_frame_range;
    p.copies_per_track = copies_per_track; p.track_stride = track_stride;
    B200X_REQUIRE(y_stride >= p.out_len, "istft: y_stride too small");
    B200X_REQUIRE(d_frame_range == nullptr || d_sumsq == nullptr, "istft: sum of squares needs the full signal");
    // hops to synthesise per copy: everything, or what the affected classifier frames read (range + 3 hops)
    const int hops = d_frame_range ? std::min(n_frames - 1, std::max(1, max_range_frames) + 3) : n_frames - 1;
    // shorter strips when there is little work per copy, so that the launch still fills the SMs (3 warm-up frames/strip)
    p.hops_per_strip = (static_cast<long long>(copies) * hops >= 30000) ? 29 : 13;
    const int strips = ceil_div(hops, p.hops_per_strip);
    dim3 grid(ceil_div(strips, DSP_WARPS), copies);
    switch (mode) {
        case 0: istft_masked_kernel<0><<<grid, DSP_THREADS, ISTFT_SMEM, s>>>(p); break;
        case 1: istft_masked_kernel<1><<<grid, DSP_THREADS, ISTFT_SMEM, s>>>(p); break;
        case 2: istft_masked_kernel<2><<<grid, DSP_THREADS, ISTFT_SMEM, s>>>(p); break;
        case 3: istft_masked_kernel<3><<<grid, DSP_THREADS, ISTFT_SMEM, s>>>(p); break;
        default: istft_masked_kernel<4><<<grid, DSP_THREADS, ISTFT_SMEM, s>>>(p); break;
    }
    B200X_CUDA_TRY(cudaGetLastError());
    return B200X_OK;
}

namespace b200x {
// HTK mel filterbank, torchaudio.functional.melscale_fbanks(norm=None, mel_scale='htk') restated; sparse rows.
struct MelBank {
    int n_mels = 0, mel_slots = 0;
    float2* d_lane_weights = nullptr;
    int2 *d_lane_info = nullptr, *d_filter_segs = nullptr;
};
struct MelBankSlot { MelBank bank; double key[3] = {0, 0, 0}; };
static std::map<int, MelBankSlot> g_banks;                 // one filterbank per DEVICE (its pointers are device allocations)
static std::mutex g_bank_mutex;

// HTK triangular filters, unnormalised (torchaudio melscale_fbanks(norm=None, mel_scale="htk")): w[f][k] =
// max(0, min(rising, falling)) in double, rounded to float.  Stored by segment (see MelParams).
static int ensure_melbank(int sample_rate, int n_mels, double f_min, double f_max, MelBank* out) {
    int dev = 0;
    B200X_CUDA_TRY(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_bank_mutex);
    MelBankSlot& slot = g_banks[dev];
    MelBank& g_bank = slot.bank;
    double* g_bank_key = slot.key;
    if (g_bank.n_mels == n_mels && g_bank_key[0] == sample_rate && g_bank_key[1] == f_min && g_bank_key[2] == f_max) {
        *out = g_bank;
        return B200X_OK;
    }
    std::vector<double> f_pts(n_mels + 2);
    const double m_min = 2595.0 * std::log10(1.0 + f_min / 700.0), m_max = 2595.0 * std::log10(1.0 + f_max / 700.0);
    for (int i = 0; i < n_mels + 2; ++i) {
        const double m = m_min + (m_max - m_min) * i / (n_mels + 1);
        f_pts[i] = 700.0 * (std::pow(10.0, m / 2595.0) - 1.0);
    }
    auto weight = [&](int f, int k) -> float {
        if (f < 0 || f >= n_mels) return 0.f;
        const double freq = static_cast<double>(k) * (sample_rate / 2) / (NBIN - 1);
        const double rising = (freq - f_pts[f]) / (f_pts[f + 1] - f_pts[f]);
        const double falling = (f_pts[f + 2] - freq) / (f_pts[f + 2] - f_pts[f + 1]);
        return static_cast<float>(std::max(0.0, std::min(rising, falling)));
    };
    // segment of bin k: the j with f_pts[j] < freq <= f_pts[j + 1] (clamped to [0, n_mels]); it feeds filters j and j - 1
    const int n_seg = n_mels + 1;
    std::vector<int> seg_of(NBIN), seg_start(n_seg + 1, NBIN);
    std::vector<float2> w(NBIN);
    for (int k = 0; k < NBIN; ++k) {
        const double freq = static_cast<double>(k) * (sample_rate / 2) / (NBIN - 1);
        int j = 0;
        while (j < n_mels && freq > f_pts[j + 1]) ++j;
        seg_of[k] = j;
        w[k] = make_float2(weight(j, k), weight(j - 1, k));
        // every other filter must be silent at this bin, or the two-filter decomposition would drop weight
        for (int f = 0; f < n_mels; ++f)
            if (f != j && f != j - 1 && weight(f, k) != 0.f)
                return set_error(B200X_ERR_INVALID, "mel bank: bin %d feeds filter %d outside its segment %d", k, f, j);
    }
    for (int k = NBIN - 1; k >= 0; --k) seg_start[seg_of[k]] = k;
    for (int j = n_seg - 1; j >= 0; --j) seg_start[j] = std::min(seg_start[j], seg_start[j + 1]);   // empty segments
    // compact the non-empty segments and cut them into 32 contiguous runs of whole segments with the smallest possible
    // longest run (binary search on the cap, greedy fill).  Bins that carry no weight at the two ends of the spectrum (below
    // f_min in the first segment, above f_max in the last one - half of the spectrum when f_max is half Nyquist) are trimmed:
    // they sit at the start of the first run / the end of the last run, so every run stays a contiguous range of bins.
    std::vector<int> eff_start(seg_start.begin(), seg_start.end() - 1), eff_end(seg_start.begin() + 1, seg_start.end());
    auto silent = [&](int k) { return w[k].x == 0.f && w[k].y == 0.f; };
    for (int j = 0; j < n_seg; ++j) {                     // leading silent bins of the spectrum
        while (eff_start[j] < eff_end[j] && silent(eff_start[j])) ++eff_start[j];
        if (eff_start[j] < eff_end[j]) break;
    }
    for (int j = n_seg - 1; j >= 0; --j) {                // trailing silent bins
        while (eff_end[j] > eff_start[j] && silent(eff_end[j] - 1)) --eff_end[j];
        if (eff_end[j] > eff_start[j]) break;
    }
    std::vector<int> compact(n_seg, -1), seg_ids;
    for (int j = 0; j < n_seg; ++j)
        if (eff_end[j] > eff_start[j]) { compact[j] = static_cast<int>(seg_ids.size()); seg_ids.push_back(j); }
    const int n_ne = static_cast<int>(seg_ids.size());
    auto width = [&](int c) { return eff_end[seg_ids[c]] - eff_start[seg_ids[c]]; };
    auto runs_needed = [&](int cap) {
        int runs = 1, fill = 0;
        for (int c = 0; c < n_ne; ++c) {
            if (width(c) > cap) return 1 << 30;
            if (fill + width(c) > cap) { ++runs; fill = 0; }
            fill += width(c);
        }
        return runs;
    };
    int lo = 1, hi = NBIN;
    while (lo < hi) {
        const int mid = (lo + hi) / 2;
        if (runs_needed(mid) <= 32) hi = mid; else lo = mid + 1;
    }
    const int cap = lo;
    const int slots = (cap + 7) / 8 * 8;
    if (slots > MEL_MAX_SLOTS) return set_error(B200X_ERR_INVALID, "mel bank: a lane would walk %d bins (limit %d)", cap, MEL_MAX_SLOTS);
    std::vector<int2> lane_info(32, make_int2(NBIN - 1, n_ne));
    std::vector<float2> lane_w(static_cast<size_t>(slots) * 32, make_float2(0.f, 0.f));
    {
        int lane = 0, fill = 0;
        for (int c = 0; c < n_ne; ++c) {
            if (fill + width(c) > cap) { ++lane; fill = 0; }
            const int j = seg_ids[c];
            if (fill == 0) lane_info[lane] = make_int2(eff_start[j], c);
            for (int k = eff_start[j]; k < eff_end[j]; ++k, ++fill) {
                float2 v = w[k];
                if (k == eff_end[j] - 1) v.x = -v.x;                 // flag: last bin of the segment (-0.0f for a zero weight)
                lane_w[static_cast<size_t>(fill) * 32 + lane] = v;
            }
        }
    }
    // contiguity check: a lane's slots must be consecutive bins (segments are, and runs are made of consecutive segments)
    std::vector<int2> filt(n_mels);
    for (int f = 0; f < n_mels; ++f) filt[f] = make_int2(compact[f], compact[f + 1]);
    if (g_bank.d_lane_weights) { cudaFree(g_bank.d_lane_weights); cudaFree(g_bank.d_lane_info); cudaFree(g_bank.d_filter_segs); }
    B200X_CUDA_TRY(cudaMalloc(&g_bank.d_lane_weights, lane_w.size() * sizeof(float2)));
    B200X_CUDA_TRY(cudaMalloc(&g_bank.d_lane_info, lane_info.size() * sizeof(int2)));
    B200X_CUDA_TRY(cudaMalloc(&g_bank.d_filter_segs, filt.size() * sizeof(int2)));
    B200X_CUDA_TRY(cudaMemcpy(g_bank.d_lane_weights, lane_w.data(), lane_w.size() * sizeof(float2), cudaMemcpyHostToDevice));
    B200X_CUDA_TRY(cudaMemcpy(g_bank.d_lane_info, lane_info.data(), lane_info.size() * sizeof(int2), cudaMemcpyHostToDevice));
    B200X_CUDA_TRY(cudaMemcpy(g_bank.d_filter_segs, filt.data(), filt.size() * sizeof(int2), cudaMemcpyHostToDevice));
    g_bank.mel_slots = slots;
    g_bank.n_mels = n_mels;
    g_bank_key[0] = sample_rate; g_bank_key[1] = f_min; g_bank_key[2] = f_max;
    *out = g_bank;
    return B200X_OK;
}
}  // namespace b200x

extern "C" int b200x_mel_frames_per_cta(void) { return 32; }

extern "C" int b200x_mel_db(const float* d_y, int64_t y_stride, int64_t n_samples, int copies, int sample_rate,
                            int n_mels, double f_min, double f_max, double amin, const double* d_sumsq,
                            double ref_rms, int64_t rms_count, float* d_db, int db_frames, float* d_cta_max,
                            const int32_t* d_frame_range, int max_range_frames, void* stream) {
    return b200x_mel_db_ref(d_y, y_stride, n_samples, copies, sample_rate, n_mels, f_min, f_max, amin, d_sumsq, ref_rms, nullptr,
                            rms_count, d_db, db_frames, d_cta_max, d_frame_range, max_range_frames, stream);
}

extern "C" int b200x_wave_rms(const float* d_waves, int64_t n_samples, int64_t stride, int n_waves, int repeat, double* d_out,
                              void* stream) {
    B200X_REQUIRE(d_waves && d_out && n_samples > 0 && n_waves > 0 && repeat > 0, "wave_rms: bad argument");
    wave_rms_kernel<<<n_waves, 1024, 0, static_cast<cudaStream_t>(stream)>>>(d_waves, n_samples, stride, repeat, d_out);
    B200X_CUDA_TRY(cudaGetLastError());
    return B200X_OK;
}

extern "C" int b200x_mel_db_ref(const float* d_y, int64_t y_stride, int64_t n_samples, int copies, int sample_rate,
                                int n_mels, double f_min, double f_max, double amin, const double* d_sumsq,
                                double ref_rms, const double* d_ref_rms_per_copy, int64_t rms_count, float* d_db, int db_frames,
                                float* d_cta_max, const int32_t* d_frame_range, int max_range_frames, void* stream) {
    B200X_REQUIRE(n_mels > 0 && n_mels <= 128 && n_mels % 32 == 0, "mel: n_mels=%d unsupported", n_mels);
    B200X_REQUIRE(n_samples > NFFT / 2 && copies > 0, "mel: bad sizes");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    B200X_TRY(ensure_tables(s));
    MelBank g_bank;
    B200X_TRY(ensure_melbank(sample_rate, n_mels, f_min, f_max, &g_bank));
    B200X_TRY(ensure_kernel_smem(reinterpret_cast<const void*>(mel_db_kernel), MEL_SMEM));
    MelParams p;
    p.y = d_y; p.y_stride = y_stride; p.n_samples = n_samples; p.sumsq = d_sumsq; p.ref_rms = ref_rms; p.ref_rms_arr = d_ref_rms_per_copy; p.rms_count = rms_count;
    p.n_frames = 1 + static_cast<int>(n_samples / HOP); p.n_mels = n_mels;
    p.lane_weights = g_bank.d_lane_weights; p.lane_info = g_bank.d_lane_info; p.filter_segs = g_bank.d_filter_segs; p.mel_slots = g_bank.mel_slots;
    p.amin = static_cast<float>(amin); p.db = d_db; p.cta_max = d_cta_max; p.frames_per_cta = b200x_mel_frames_per_cta();
    p.db_frames = db_frames; p.frame_range = d_frame_range;
    const int span = d_frame_range ? std::min(p.n_frames, std::max(1, max_range_frames)) : p.n_frames;
    B200X_REQUIRE(db_frames >= span, "mel: db_frames=%d smaller than the frame span %d", db_frames, span);
    dim3 grid(ceil_div(span, p.frames_per_cta), copies);
    mel_db_kernel<<<grid, DSP_THREADS, MEL_SMEM, s>>>(p);
    B200X_CUDA_TRY(cudaGetLastError());
    return B200X_OK;
}

extern "C" int b200x_mel_normalize_resize(const float* d_db, int db_frames, const float* d_cta_max, int n_cta_max, int copies,
                                          int n_frames, int n_mels, float top_db, int unbiased, float eps, int out_t,
                                          const float* d_db_base, const float* d_base_premax, const float* d_base_sufmax,
                                          const int32_t* d_frame_range, void* d_partial, float* d_floor, void* d_img_t,
                                          void* d_img_f, int ld_f, void* stream) {
    B200X_REQUIRE(n_mels <= 128 && n_mels % 4 == 0, "resize: bad sizes");
    B200X_REQUIRE(d_db_base == nullptr || (d_base_premax && d_base_sufmax && d_frame_range), "resize: baseline needs maxima and frame ranges");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int n_partial = 32;
    StatsParams sp;
    sp.db = d_db; sp.db_frames = db_frames; sp.n_frames = n_frames; sp.n_mels = n_mels; sp.cta_max = d_cta_max;
    sp.n_cta_max = n_cta_max; sp.top_db = top_db; sp.base = d_db_base; sp.base_premax = d_base_premax;
    sp.base_sufmax = d_base_sufmax; sp.frame_range = d_frame_range; sp.partial = reinterpret_cast<double2*>(d_partial);
    sp.floor_out = d_floor;
    dim3 g1(n_partial, copies);
    mel_stats_kernel<<<g1, 256, 0, s>>>(sp);
    B200X_CUDA_TRY(cudaGetLastError());
    ResizeParams p;
    p.db = d_db; p.db_frames = db_frames; p.base = d_db_base; p.frame_range = d_frame_range;
    p.partial = reinterpret_cast<const double2*>(d_partial); p.floor_val = d_floor; p.n_partial = n_partial;
    p.n_frames = n_frames; p.n_mels = n_mels; p.out_t = out_t; p.unbiased = unbiased; p.eps = eps;
    p.img_t = reinterpret_cast<__nv_bfloat16*>(d_img_t); p.img_f = reinterpret_cast<__nv_bfloat16*>(d_img_f); p.ld_f = ld_f;
    dim3 g2(ceil_div(out_t, 64), copies);
    mel_resize_kernel<<<g2, 256, 0, s>>>(p);
    B200X_CUDA_TRY(cudaGetLastError());
    return B200X_OK;
}

extern "C" int b200x_frame_ranges(const int32_t* d_windows, int n, int n_frames, int32_t* d_ranges, void* stream) {
    if (n <= 0) return B200X_OK;
    frame_ranges_kernel<<<ceil_div(n, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(d_windows, n, n_frames, d_ranges);
    B200X_CUDA_TRY(cudaGetLastError());
    return B200X_OK;
}

extern "C" int b200x_mel_base_maxima(const float* d_db_base, int n_frames, int n_mels, float* d_premax, float* d_sufmax,
                                     void* stream) {
    B200X_REQUIRE(n_frames > 0 && n_frames <= 12000, "base_maxima: n_frames=%d out of range", n_frames);
    base_maxima_kernel<<<1, 1024, n_frames * sizeof(float), static_cast<cudaStream_t>(stream)>>>(d_db_base, n_frames, n_mels, d_premax, d_sufmax);
    B200X_CUDA_TRY(cudaGetLastError());
    return B200X_OK;
}

extern "C" int b200x_mix_stems(const float* d_stems, int64_t n_samples, int n_stems, const uint8_t* d_masks, int copies,
                               float* d_y, int64_t y_stride, void* stream) {
    B200X_REQUIRE(n_stems > 0 && copies > 0 && n_samples > 0, "mix_stems: bad sizes");
    dim3 grid(296, copies);
    mix_stems_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_stems, n_samples, n_stems, d_masks, d_y, y_stride);
    B200X_CUDA_TRY(cudaGetLastError());
    return B200X_OK;
}
