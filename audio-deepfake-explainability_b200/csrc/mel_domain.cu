// Mel-domain explainer variant (spec_type: mel) on sm_100a: the kernels around the existing STFT / iSTFT ones.
//
// Reference: src/spectrogram_explainability.py:367-377 (librosa.feature.melspectrogram) and :394-402
// (librosa.feature.inverse.mel_to_audio = NNLS mel -> |STFT| followed by Griffin-Lim with unseeded random phase).  The
// inverse is not reproducible even reference-vs-reference, so the arithmetic is BUILDER-DEFINED (oracle/mel.py restates it
// for the CPU): clipped least squares + projected-gradient NNLS, hashed initial phases, librosa's fast Griffin-Lim recursion.
//   mel_power_kernel : M[t][i] = sum_k A[i][k] |S[t][k]|^2                              (once per track)
//   mel_nnls_kernel  : per perturbed copy and frame: B = masked M[t] -> X = max(0, A^+ B) -> nnls_iter steps of
//                      X <- max(0, X - s A^T (A X - B)) -> |STFT| = sqrt(X).  The perturbation (occlusion rectangle over mel
//                      bins, keep-only rectangle, per-mel-bin band gain) is applied while B is loaded, so the perturbed mel
//                      spectrograms are never materialised.
//   gl_init_kernel   : C = mag * exp(2 pi i u(seed, index, cell))
//   gl_update_kernel : a = R - c T;  C = mag * a / (|a| + tiny)      (R = STFT(iSTFT(C)) of this iteration, T of the previous)
// All four are HBM / L2-bound streaming kernels; the filterbank is sparse (every FFT bin feeds at most two adjacent filters).
#include <algorithm>
#include <cmath>
#include <vector>

#include "common.h"

namespace b200x {

constexpr int MD_NBIN = 1025;
constexpr int MD_STRIDE = 1028;       // row stride of spectra / magnitudes (complex or float elements)
constexpr int MD_FR = 4;              // frames per CTA of the NNLS kernel
constexpr int MD_THREADS = 256;
constexpr int MD_KPT = (MD_NBIN + MD_THREADS - 1) / MD_THREADS;    // bins per thread (5)

__global__ void __launch_bounds__(MD_THREADS)
mel_power_kernel(const float2* __restrict__ S, int stride, int n_frames, int n_mels, const float* __restrict__ basis,
                 const int* __restrict__ bin_range, float* __restrict__ mel) {
    __shared__ float pw[MD_NBIN];
    const int t = blockIdx.x;
    const float2* row = S + static_cast<long long>(t) * stride;
    for (int k = threadIdx.x; k < MD_NBIN; k += MD_THREADS) {
        const float2 z = row[k];
        pw[k] = z.x * z.x + z.y * z.y;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_mels; i += MD_THREADS) {
        const int lo = bin_range[2 * i], hi = bin_range[2 * i + 1];
        const float* a = basis + static_cast<long long>(i) * MD_NBIN;
        float acc = 0.f;
        for (int k = lo; k < hi; ++k) acc = fmaf(__ldg(a + k), pw[k], acc);
        mel[static_cast<long long>(t) * n_mels + i] = acc;
    }
}

struct NnlsParams {
    const float* mel;          // [n_frames][n_mels] power mel spectrogram of the track
    int n_frames, n_mels;
    int mode;                  // B200X_MASK_NONE / OCCLUDE / BAND_GAIN / KEEP_ONLY (over mel bins)
    const int* windows;        // [copies][4] = t0, t1, f0 (mel), f1 (mel)
    float occlusion_value;
    const float* gains;        // [copies][n_mels]
    const float* basis;        // dense [n_mels][1025]
    const int* bin_range;      // [n_mels][2]: bins with a non-zero weight
    const int* bin_first;      // [1025]: first filter fed by bin k
    const float2* bin_w;       // [1025]: weights into filters first, first + 1
    const float* pinv_t;       // [n_mels][1025]: rows of pinv(A)^T
    float step;                // 1 / ||A||_2^2
    int nnls_iter;
    const int* frame_range;    // [copies][2]: frames [fa, fb) to solve for
    float* mag;                // [copies][mag_copy_stride] rows of MD_STRIDE floats
    long long mag_copy_stride;
};

__global__ void __launch_bounds__(MD_THREADS)
mel_nnls_kernel(NnlsParams p) {
    extern __shared__ float md_smem[];
    float* Bs = md_smem;                               // [FR][n_mels]
    float* Rs = Bs + MD_FR * p.n_mels;                 // [FR][n_mels]
    float* Xs = Rs + MD_FR * p.n_mels;                 // [FR][MD_STRIDE]
    const int copy = blockIdx.y;
    const int fa = p.frame_range[2 * copy], fb = p.frame_range[2 * copy + 1];
    const int frame0 = fa + blockIdx.x * MD_FR;
    if (frame0 >= fb) return;
    int t0 = 0, t1 = 0, f0 = 0, f1 = 0;
    if (p.mode == B200X_MASK_OCCLUDE || p.mode == B200X_MASK_KEEP_ONLY) {
        const int4 w = *reinterpret_cast<const int4*>(p.windows + 4 * copy);
        t0 = w.x; t1 = w.y; f0 = w.z; f1 = w.w;
    }
    const float* gain = p.mode == B200X_MASK_BAND_GAIN ? p.gains + static_cast<long long>(copy) * p.n_mels : nullptr;
    // ---- B: the perturbed mel column of every frame of this CTA (the perturbation lives here and nowhere else)
    for (int idx = threadIdx.x; idx < MD_FR * p.n_mels; idx += MD_THREADS) {
        const int f = idx / p.n_mels, i = idx - f * p.n_mels;
        const int t = frame0 + f;
        float v = (t < fb && t < p.n_frames) ? p.mel[static_cast<long long>(t) * p.n_mels + i] : 0.f;
        const bool inside = t >= t0 && t < t1 && i >= f0 && i < f1;
        if (p.mode == B200X_MASK_OCCLUDE) { if (inside) v = p.occlusion_value; }
        else if (p.mode == B200X_MASK_KEEP_ONLY) { if (!inside) v = 0.f; }
        else if (p.mode == B200X_MASK_BAND_GAIN) v *= gain[i];
        Bs[idx] = v;
    }
    __syncthreads();
    // ---- X0 = max(0, pinv(A) B): every thread owns bins tid + 256 m of all FR frames
    float acc[MD_FR][MD_KPT];
#pragma unroll
    for (int f = 0; f < MD_FR; ++f)
#pragma unroll
        for (int m = 0; m < MD_KPT; ++m) acc[f][m] = 0.f;
    for (int i = 0; i < p.n_mels; ++i) {
        const float* prow = p.pinv_t + static_cast<long long>(i) * MD_NBIN;
        float b[MD_FR];
#pragma unroll
        for (int f = 0; f < MD_FR; ++f) b[f] = Bs[f * p.n_mels + i];
#pragma unroll
        for (int m = 0; m < MD_KPT; ++m) {
            const int k = threadIdx.x + MD_THREADS * m;
            const float pv = k < MD_NBIN ? __ldg(prow + k) : 0.f;
#pragma unroll
            for (int f = 0; f < MD_FR; ++f) acc[f][m] = fmaf(pv, b[f], acc[f][m]);
        }
    }
#pragma unroll
    for (int f = 0; f < MD_FR; ++f)
#pragma unroll
        for (int m = 0; m < MD_KPT; ++m) {
            const int k = threadIdx.x + MD_THREADS * m;
            if (k < MD_STRIDE) Xs[f * MD_STRIDE + k] = k < MD_NBIN ? fmaxf(acc[f][m], 0.f) : 0.f;
        }
    __syncthreads();
    // ---- projected gradient: R = A X - B (gather over each filter's bins), X <- max(0, X - s A^T R) (two filters per bin)
    for (int it = 0; it < p.nnls_iter; ++it) {
        for (int idx = threadIdx.x; idx < MD_FR * p.n_mels; idx += MD_THREADS) {
            const int f = idx / p.n_mels, i = idx - f * p.n_mels;
            const int lo = p.bin_range[2 * i], hi = p.bin_range[2 * i + 1];
            const float* a = p.basis + static_cast<long long>(i) * MD_NBIN;
            const float* x = Xs + f * MD_STRIDE;
            float s = 0.f;
            for (int k = lo; k < hi; ++k) s = fmaf(__ldg(a + k), x[k], s);
            Rs[idx] = s - Bs[idx];
        }
        __syncthreads();
#pragma unroll
        for (int m = 0; m < MD_KPT; ++m) {
            const int k = threadIdx.x + MD_THREADS * m;
            if (k < MD_NBIN) {
                const int i0 = p.bin_first[k];
                const float2 w = p.bin_w[k];
#pragma unroll
                for (int f = 0; f < MD_FR; ++f) {
                    const float* r = Rs + f * p.n_mels;
                    float g = w.x * r[i0];
                    if (i0 + 1 < p.n_mels) g = fmaf(w.y, r[i0 + 1], g);
                    Xs[f * MD_STRIDE + k] = fmaxf(Xs[f * MD_STRIDE + k] - p.step * g, 0.f);
                }
            }
        }
        __syncthreads();
    }
    // ---- |STFT| = X^(1/2)  (mel_to_stft, power = 2)
    for (int f = 0; f < MD_FR; ++f) {
        const int t = frame0 + f;
        if (t >= fb || t >= p.n_frames) break;
        float* out = p.mag + static_cast<long long>(copy) * p.mag_copy_stride + static_cast<long long>(t) * MD_STRIDE;
        for (int k = threadIdx.x; k < MD_STRIDE; k += MD_THREADS) out[k] = k < MD_NBIN ? sqrtf(Xs[f * MD_STRIDE + k]) : 0.f;
    }
}

// C = mag * e^(2 pi i u), u = hash(key(seed, index) ^ cell) / 2^32 (cell = frame * 1025 + bin: the RISE hash, common.h)
__global__ void gl_init_kernel(const float* __restrict__ mag, long long mag_copy_stride, float2* __restrict__ C,
                               long long c_copy_stride, int n_frames, uint32_t seed, int first_index) {
    const int copy = blockIdx.y;
    const uint32_t key = rise_mask_key(seed, static_cast<uint32_t>(first_index + copy));
    const long long total = static_cast<long long>(n_frames) * MD_STRIDE;
    const float* m = mag + static_cast<long long>(copy) * mag_copy_stride;
    float2* c = C + static_cast<long long>(copy) * c_copy_stride;
    for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int t = static_cast<int>(e / MD_STRIDE), k = static_cast<int>(e - static_cast<long long>(t) * MD_STRIDE);
        float2 out = make_float2(0.f, 0.f);
        if (k < MD_NBIN) {
            const uint32_t h = hash_lowbias32(key ^ ((static_cast<uint32_t>(t) * MD_NBIN + k) * 0xC2B2AE35U));
            const float u = __uint2float_rn(h) * 2.3283064365386963e-10f;      // h / 2^32, rounded to float32 like the oracle
            float sn, cs;
            sincospif(2.0f * u, &sn, &cs);
            const float a = m[e];
            out = make_float2(a * cs, a * sn);
        }
        c[e] = out;
    }
}

// librosa.griffinlim inner update: angles = rebuilt - coef * tprev; angles /= |angles| + tiny; angles *= S
__global__ void gl_update_kernel(const float2* __restrict__ R, const float2* __restrict__ T, const float* __restrict__ mag,
                                 long long mag_copy_stride, float2* __restrict__ C, long long c_copy_stride, int n_frames, float coef) {
    const int copy = blockIdx.y;
    const long long total = static_cast<long long>(n_frames) * MD_STRIDE;
    const long long off = static_cast<long long>(copy) * c_copy_stride;
    const float* m = mag + static_cast<long long>(copy) * mag_copy_stride;
    for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float2 r = R[off + e];
        float2 a = r;
        if (coef != 0.f) {
            const float2 tp = T[off + e];
            a = make_float2(r.x - coef * tp.x, r.y - coef * tp.y);
        }
        const float inv = 1.0f / (sqrtf(a.x * a.x + a.y * a.y) + 1.17549435e-38f);
        const float g = m[e];
        C[off + e] = make_float2(g * a.x * inv, g * a.y * inv);
    }
}

}  // namespace b200x

using namespace b200x;

extern "C" int b200x_mel_power(const void* d_spec, int spec_stride, int n_frames, int n_mels, const float* d_basis,
                               const int32_t* d_bin_range, float* d_mel, void* stream) {
    B200X_REQUIRE(d_spec && d_basis && d_bin_range && d_mel, "mel_power: NULL argument");
    B200X_REQUIRE(n_frames > 0 && n_mels > 0 && spec_stride >= MD_NBIN, "mel_power: bad sizes");
    mel_power_kernel<<<n_frames, MD_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const float2*>(d_spec), spec_stride, n_frames,
                                                                                    n_mels, d_basis, d_bin_range, d_mel);
    B200X_CUDA_TRY(cudaGetLastError());
    return B200X_OK;
}

extern "C" int b200x_mel_nnls(const float* d_mel, int n_frames, int n_mels, int copies, int mode, const int32_t* d_windows,
                              float occlusion_value, const float* d_gains, const float* d_basis, const int32_t* d_bin_range,
                              const int32_t* d_bin_first, const float* d_bin_w, const float* d_pinv_t, float step, int nnls_iter,
                              const int32_t* d_frame_range, int max_range_frames, float* d_mag, int64_t mag_copy_stride,
                              void* stream) {
    B200X_REQUIRE(d_mel && d_basis && d_bin_range && d_bin_first && d_bin_w && d_pinv_t && d_frame_range && d_mag, "mel_nnls: NULL argument");
    B200X_REQUIRE(n_frames > 0 && n_mels > 0 && n_mels <= 1024 && copies > 0 && nnls_iter >= 0 && max_range_frames > 0, "mel_nnls: bad sizes");
    B200X_REQUIRE(mode == B200X_MASK_NONE || mode == B200X_MASK_OCCLUDE || mode == B200X_MASK_BAND_GAIN || mode == B200X_MASK_KEEP_ONLY,
                  "mel_nnls: bad mode %d", mode);
    B200X_REQUIRE((mode != B200X_MASK_OCCLUDE && mode != B200X_MASK_KEEP_ONLY) || d_windows != nullptr, "mel_nnls: windows missing");
    B200X_REQUIRE(mode != B200X_MASK_BAND_GAIN || d_gains != nullptr, "mel_nnls: gains missing");
    B200X_REQUIRE(mag_copy_stride >= static_cast<int64_t>(n_frames) * MD_STRIDE, "mel_nnls: mag_copy_stride too small");
    NnlsParams p;
    p.mel = d_mel; p.n_frames = n_frames; p.n_mels = n_mels; p.mode = mode; p.windows = d_windows; p.occlusion_value = occlusion_value;
    p.gains = d_gains; p.basis = d_basis; p.bin_range = d_bin_range; p.bin_first = d_bin_first;
    p.bin_w = reinterpret_cast<const float2*>(d_bin_w); p.pinv_t = d_pinv_t; p.step = step; p.nnls_iter = nnls_iter;
    p.frame_range = d_frame_range; p.mag = d_mag; p.mag_copy_stride = mag_copy_stride;
    const int smem = (2 * MD_FR * n_mels + MD_FR * MD_STRIDE) * static_cast<int>(sizeof(float));
    B200X_TRY(ensure_kernel_smem(reinterpret_cast<const void*>(mel_nnls_kernel), 2 * MD_FR * 1024 * 4 + MD_FR * MD_STRIDE * 4));
    dim3 grid(ceil_div(max_range_frames, MD_FR), copies);
    mel_nnls_kernel<<<grid, MD_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(p);
    B200X_CUDA_TRY(cudaGetLastError());
    return B200X_OK;
}

extern "C" int b200x_gl_init(const float* d_mag, int64_t mag_copy_stride, void* d_c, int64_t c_copy_stride, int copies, int n_frames,
                             uint32_t seed, int first_index, void* stream) {
    B200X_REQUIRE(d_mag && d_c && copies > 0 && n_frames > 0 && first_index >= 0, "gl_init: bad argument");
    dim3 grid(std::min<long long>(1184, (static_cast<long long>(n_frames) * MD_STRIDE + 255) / 256), copies);
    gl_init_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_mag, mag_copy_stride, reinterpret_cast<float2*>(d_c), c_copy_stride, n_frames,
                                                                       seed, first_index);
    B200X_CUDA_TRY(cudaGetLastError());
    return B200X_OK;
}

extern "C" int b200x_gl_update(const void* d_rebuilt, const void* d_tprev, const float* d_mag, int64_t mag_copy_stride, void* d_c,
                               int64_t c_copy_stride, int copies, int n_frames, float coef, void* stream) {
    B200X_REQUIRE(d_rebuilt && d_mag && d_c && copies > 0 && n_frames > 0, "gl_update: bad argument");
    B200X_REQUIRE(coef == 0.f || d_tprev != nullptr, "gl_update: previous spectrum missing");
    dim3 grid(std::min<long long>(1184, (static_cast<long long>(n_frames) * MD_STRIDE + 255) / 256), copies);
    gl_update_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const float2*>(d_rebuilt), reinterpret_cast<const float2*>(d_tprev),
                                                                         d_mag, mag_copy_stride, reinterpret_cast<float2*>(d_c), c_copy_stride, n_frames, coef);
    B200X_CUDA_TRY(cudaGetLastError());
    return B200X_OK;
}
