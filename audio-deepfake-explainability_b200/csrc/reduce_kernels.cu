// Row-wise and reduction kernels of the hot path (HBM/L2-bound; warp-shuffle reductions, 16-byte accesses):
//   * layernorm_kernel ........ LayerNorm over the embedding dim, fp32 in -> bf16 (GEMM operand) or fp32 in place
//   * head_partial/finish ..... final LayerNorm + token mean + Linear(D,1) + sigmoid -> fake-probability per copy
//   * saliency_kernel ......... delta-prob -> importance map (gather form of map[f0:f1,t0:t1] += d; count += 1;
//                               map /= count + 1e-8, src/spectrogram_explainability.py:695-707), float64, window order
//   * band_map_kernel ......... FBP rows: map[(freqs>=low)&(freqs<=high), :] += delta (src/dsp_band_ops.py:652-653)
//   * rank_kernel ............. stable ranking of windows by key (Python sorted() semantics incl. ties)
#include "common.h"
#include "layernorm_rows.cuh"

namespace b200x {

// one warp per row; D <= 1024, D % 128 == 0 handled with float4 lanes (D = 384 -> 3 float4 per lane)
template <int VPL>   // float4 vectors per lane
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, int rows, int D, const float* __restrict__ gamma_a,
                 const float* __restrict__ beta_a, const float* __restrict__ gamma_b, const float* __restrict__ beta_b,
                 int group, int split, float eps, __nv_bfloat16* __restrict__ out_bf16, float* __restrict__ out_f32, int reverse) {
    const int blk = reverse ? static_cast<int>(gridDim.x) - 1 - static_cast<int>(blockIdx.x) : static_cast<int>(blockIdx.x);
    const int row = blk * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    // rows are grouped (group rows per copy); rows with (row % group) >= split use the second parameter set
    const bool second = group > 0 && (row % group) >= split;
    const float* gamma = second ? gamma_b : gamma_a;
    const float* beta = second ? beta_b : beta_a;
    const float4* src = reinterpret_cast<const float4*>(x + static_cast<long long>(row) * D);
    if constexpr (VPL <= 3) {
        // widths up to 384: the shared row routine (the LayerNorm tail of the residual GEMM runs the same code)
        float4 v[1][3];
#pragma unroll
        for (int i = 0; i < 3; ++i) v[0][i] = i < VPL ? src[lane + 32 * i] : make_float4(0.f, 0.f, 0.f, 0.f);
        layernorm_rows<1>(v, VPL, 1.0f / static_cast<float>(D), eps, gamma, beta, lane, [&](int, int i, float4 o) {
            if (out_bf16 != nullptr)
                reinterpret_cast<uint2*>(out_bf16 + static_cast<long long>(row) * D)[lane + 32 * i] = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
            else
                reinterpret_cast<float4*>(out_f32 + static_cast<long long>(row) * D)[lane + 32 * i] = o;
        });
        return;
    }
    float4 v[VPL];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        v[i] = src[lane + 32 * i];
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
        q += (a * a + b * b) + (c * c + d * d);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q / D + eps);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const float4 g = reinterpret_cast<const float4*>(gamma)[lane + 32 * i];
        const float4 b = reinterpret_cast<const float4*>(beta)[lane + 32 * i];
        float4 o;
        o.x = (v[i].x - mean) * rstd * g.x + b.x;
        o.y = (v[i].y - mean) * rstd * g.y + b.y;
        o.z = (v[i].z - mean) * rstd * g.z + b.z;
        o.w = (v[i].w - mean) * rstd * g.w + b.w;
        if (out_bf16 != nullptr) {
            reinterpret_cast<uint2*>(out_bf16 + static_cast<long long>(row) * D)[lane + 32 * i] = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
        } else {
            reinterpret_cast<float4*>(out_f32 + static_cast<long long>(row) * D)[lane + 32 * i] = o;
        }
    }
}

// partial[copy][slice] = sum over the slice's tokens of dot(LN(x_token), w)
template <int VPL>
__global__ void __launch_bounds__(256)
head_partial_kernel(const float* __restrict__ x, int tokens, int D, const float* __restrict__ gamma,
                    const float* __restrict__ beta, float eps, int use_norm, const float* __restrict__ w,
                    float* __restrict__ partial) {
    __shared__ float s_part[8];
    const int copy = blockIdx.y, slice = blockIdx.x, n_slices = gridDim.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int per = (tokens + n_slices - 1) / n_slices;
    const int t_begin = slice * per, t_end = min(t_begin + per, tokens);
    float acc = 0.f;
    for (int t = t_begin + warp; t < t_end; t += 8) {
        const float4* src = reinterpret_cast<const float4*>(x + (static_cast<long long>(copy) * tokens + t) * D);
        float4 v[VPL];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            v[i] = src[lane + 32 * i];
            s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
        float mean = 0.f, rstd = 1.f;
        if (use_norm) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            mean = s / D;
            float q = 0.f;
#pragma unroll
            for (int i = 0; i < VPL; ++i) {
                const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
                q += (a * a + b * b) + (c * c + d * d);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
            rstd = rsqrtf(q / D + eps);
        }
        float dot = 0.f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const float4 ww = reinterpret_cast<const float4*>(w)[lane + 32 * i];
            float4 o = v[i];
            if (use_norm) {
                const float4 g = reinterpret_cast<const float4*>(gamma)[lane + 32 * i];
                const float4 b = reinterpret_cast<const float4*>(beta)[lane + 32 * i];
                o.x = (o.x - mean) * rstd * g.x + b.x; o.y = (o.y - mean) * rstd * g.y + b.y;
                o.z = (o.z - mean) * rstd * g.z + b.z; o.w = (o.w - mean) * rstd * g.w + b.w;
            }
            dot += (o.x * ww.x + o.y * ww.y) + (o.z * ww.z + o.w * ww.w);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
        acc += dot;
    }
    if (lane == 0) s_part[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += s_part[i];
        partial[copy * n_slices + slice] = t;
    }
}

__global__ void head_finish_kernel(const float* __restrict__ partial, int n_slices, int tokens, float bias, int copies,
                                   float* __restrict__ logit, float* __restrict__ prob) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= copies) return;
    float s = 0.f;
    for (int i = 0; i < n_slices; ++i) s += partial[c * n_slices + i];
    const float z = s / tokens + bias;
    if (logit != nullptr) logit[c] = z;
    prob[c] = 1.0f / (1.0f + expf(-z));
}

// importance map, gather form.  Each CTA owns a 16 x 128 (freq x time) tile, compacts the windows that touch it into
// shared memory (window order preserved) and every cell adds its covering windows' deltas in that order in float64.
__global__ void __launch_bounds__(256)
saliency_kernel(const int* __restrict__ windows, const double* __restrict__ delta, int n_windows, int n_freq, int n_time,
                double* __restrict__ map) {
    extern __shared__ int s_idx[];
    __shared__ int s_n;
    const int t_lo = blockIdx.x * 128, f_lo = blockIdx.y * 16;
    const int t_hi = min(t_lo + 128, n_time), f_hi = min(f_lo + 16, n_freq);
    if (threadIdx.x == 0) {
        int n = 0;
        for (int i = 0; i < n_windows; ++i) {
            const int4 w = *reinterpret_cast<const int4*>(windows + 4 * i);
            if (w.x < t_hi && w.y > t_lo && w.z < f_hi && w.w > f_lo) s_idx[n++] = i;
        }
        s_n = n;
    }
    __syncthreads();
    const int n = s_n;
    for (int cell = threadIdx.x; cell < 16 * 128; cell += blockDim.x) {
        const int f = f_lo + cell / 128, t = t_lo + cell % 128;
        if (f >= n_freq || t >= n_time) continue;
        double acc = 0.0, cnt = 0.0;
        for (int j = 0; j < n; ++j) {
            const int i = s_idx[j];
            const int4 w = *reinterpret_cast<const int4*>(windows + 4 * i);
            if (t >= w.x && t < w.y && f >= w.z && f < w.w) { acc += delta[i]; cnt += 1.0; }
        }
        map[static_cast<long long>(f) * n_time + t] = acc / (cnt + 1e-8);
    }
}

__global__ void band_map_kernel(const int* __restrict__ rows, const double* __restrict__ delta, int n_bands, int n_freq,
                                int n_time, double* __restrict__ map) {
    const int f = blockIdx.y;
    double acc = 0.0;
    for (int b = 0; b < n_bands; ++b)
        if (f >= rows[2 * b] && f < rows[2 * b + 1]) acc += delta[b];
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n_time; t += gridDim.x * blockDim.x)
        map[static_cast<long long>(f) * n_time + t] = acc;
}

// order[rank] = index; rank_i = #{j : key_j before key_i} with ties broken by original index (stable).
// mode 0: |v| descending, 1: |v| ascending, 2: v descending, 3: v ascending
__global__ void rank_kernel(const double* __restrict__ v, int n, int mode, int* __restrict__ order) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool use_abs = mode < 2, desc = (mode == 0 || mode == 2);
    const double ki = use_abs ? fabs(v[i]) : v[i];
    int rank = 0;
    for (int j = 0; j < n; ++j) {
        const double kj = use_abs ? fabs(v[j]) : v[j];
        const bool before = desc ? (kj > ki) : (kj < ki);
        rank += (before || (kj == ki && j < i)) ? 1 : 0;
    }
    order[rank] = i;
}

__global__ void delta_kernel(const float* __restrict__ prob, float baseline, int n, double* __restrict__ delta) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) delta[i] = static_cast<double>(baseline) - static_cast<double>(prob[i]);
}

// Rational polyphase resampler (the track loader's resampling step): y[n] = up * sum_k h[n * down - k * up + c] x[k]
// with c = (h_len - 1) / 2 (zero-phase FIR, odd length) - scipy.signal.resample_poly's definition with a user filter.  One
// output sample per thread; the ~h_len / up taps of a phase are accumulated in float64 (the host path does the same).
__global__ void resample_poly_kernel(const float* __restrict__ x, long long n_in, int up, int down, const double* __restrict__ h, int h_len,
                                     float* __restrict__ y, long long n_out) {
    const long long n = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (n >= n_out) return;
    const long long c = (h_len - 1) / 2;
    const long long pos = n * down + c;                    // tap index that multiplies x[0]
    // taps j = pos - k * up must lie in [0, h_len): k from ceil((pos - h_len + 1) / up) to floor(pos / up)
    long long k_hi = pos / up;
    long long k_lo = pos - (h_len - 1);
    k_lo = k_lo <= 0 ? 0 : (k_lo + up - 1) / up;
    if (k_hi > n_in - 1) k_hi = n_in - 1;
    double acc = 0.0;
    for (long long k = k_lo; k <= k_hi; ++k) acc = fma(__ldg(h + (pos - k * up)), static_cast<double>(__ldg(x + k)), acc);
    y[n] = static_cast<float>(acc * up);
}

__global__ void delta_dev_kernel(const float* __restrict__ prob, const float* __restrict__ baseline, int n, double* __restrict__ delta) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) delta[i] = static_cast<double>(__ldg(baseline)) - static_cast<double>(prob[i]);
}

}  // namespace b200x

using namespace b200x;

extern "C" int b200x_layernorm(const float* d_x, int rows, int dim, const float* d_gamma, const float* d_beta,
                               const float* d_gamma2, const float* d_beta2, int group, int split, float eps,
                               void* d_out_bf16, float* d_out_f32, int reverse, void* stream) {
    B200X_REQUIRE(rows > 0 && dim % 128 == 0 && dim <= 1024, "layernorm: dim=%d must be a multiple of 128 (<= 1024)", dim);
    B200X_REQUIRE((d_out_bf16 != nullptr) != (d_out_f32 != nullptr), "layernorm: exactly one output");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int grid = ceil_div(rows, 8);
    __nv_bfloat16* ob = reinterpret_cast<__nv_bfloat16*>(d_out_bf16);
#define LN_CASE(V) case V: layernorm_kernel<V><<<grid, 256, 0, s>>>(d_x, rows, dim, d_gamma, d_beta, d_gamma2 ? d_gamma2 : d_gamma, d_beta2 ? d_beta2 : d_beta, group, split, eps, ob, d_out_f32, reverse != 0); break;
    switch (dim / 128) { LN_CASE(1) LN_CASE(2) LN_CASE(3) LN_CASE(4) LN_CASE(6) LN_CASE(8)
        default: return set_error(B200X_ERR_INVALID, "layernorm: dim=%d not instantiated", dim); }
#undef LN_CASE
    B200X_CUDA_TRY(cudaGetLastError());
    return B200X_OK;
}

extern "C" int b200x_head(const float* d_x, int copies, int tokens, int dim, const float* d_gamma, const float* d_beta,
                          float eps, int use_norm, const float* d_w, float bias, float* d_partial, float* d_logit,
                          float* d_prob, void* stream) {
    B200X_REQUIRE(copies > 0 && tokens > 0 && dim % 128 == 0 && dim <= 1024, "head: bad sizes");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int n_slices = 8;
    dim3 grid(n_slices, copies);
#define HD_CASE(V) case V: head_partial_kernel<V><<<grid, 256, 0, s>>>(d_x, tokens, dim, d_gamma, d_beta, eps, use_norm, d_w, d_partial); break;
    switch (dim / 128) { HD_CASE(1) HD_CASE(2) HD_CASE(3) HD_CASE(4) HD_CASE(6) HD_CASE(8)
        default: return set_error(B200X_ERR_INVALID, "head: dim=%d not instantiated", dim); }
#undef HD_CASE
    B200X_CUDA_TRY(cudaGetLastError());
    head_finish_kernel<<<ceil_div(copies, 128), 128, 0, s>>>(d_partial, n_slices, tokens, bias, copies, d_logit, d_prob);
    B200X_CUDA_TRY(cudaGetLastError());
    return B200X_OK;
}

extern "C" int b200x_head_slices(void) { return 8; }

extern "C" int b200x_resample_poly(const float* d_x, int64_t n_in, int up, int down, const double* d_h, int h_len, float* d_y,
                                   int64_t n_out, void* stream) {
    B200X_REQUIRE(d_x && d_h && d_y && n_in > 0 && n_out > 0 && up > 0 && down > 0 && h_len > 0 && (h_len & 1), "resample_poly: bad argument");
    resample_poly_kernel<<<static_cast<unsigned>((n_out + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(d_x, n_in, up, down, d_h, h_len, d_y, n_out);
    B200X_CUDA_TRY(cudaGetLastError());
    return B200X_OK;
}

extern "C" int b200x_delta_dev(const float* d_prob, const float* d_baseline, int n, double* d_delta, void* stream) {
    if (n <= 0) return B200X_OK;
    B200X_REQUIRE(d_prob && d_baseline && d_delta, "delta_dev: NULL argument");
    delta_dev_kernel<<<ceil_div(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(d_prob, d_baseline, n, d_delta);
    B200X_CUDA_TRY(cudaGetLastError());
    return B200X_OK;
}

extern "C" int b200x_delta(const float* d_prob, float baseline, int n, double* d_delta, void* stream) {
    if (n <= 0) return B200X_OK;
    delta_kernel<<<ceil_div(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(d_prob, baseline, n, d_delta);
    B200X_CUDA_TRY(cudaGetLastError());
    return B200X_OK;
}

namespace b200x {
// RISE accumulation (src/spectrogram_explainability.py:783, :798): map[f][t] = sum_i keep_i(f, t) * pred[i], float64 in mask
// order (adding the reference's 0.0 * pred terms would not change a bit), then / (n_masks * p + 1e-8).  The keep bits are
// re-derived from the hash, so no mask is ever stored.
__global__ void __launch_bounds__(256)
rise_map_kernel(const double* __restrict__ pred, int n_masks, uint32_t seed, uint32_t threshold, double denom, int n_freq,
                int n_time, double* __restrict__ map) {
    extern __shared__ uint32_t s_keys[];
    for (int i = threadIdx.x; i < n_masks; i += blockDim.x) s_keys[i] = rise_mask_key(seed, static_cast<uint32_t>(i));
    __syncthreads();
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int f = blockIdx.y;
    if (t >= n_time) return;
    const uint32_t cell = static_cast<uint32_t>(t) * static_cast<uint32_t>(n_freq) + static_cast<uint32_t>(f);
    double acc = 0.0;
    for (int i = 0; i < n_masks; ++i)
        if (rise_keep(s_keys[i], cell, threshold)) acc += pred[i];
    map[static_cast<long long>(f) * n_time + t] = acc / denom;
}
}  // namespace b200x

extern "C" int b200x_rise_map(const double* d_pred, int n_masks, uint32_t seed, double keep_probability, int n_freq, int n_time,
                              double* d_map, void* stream) {
    B200X_REQUIRE(n_freq > 0 && n_time > 0 && n_masks >= 0 && n_masks <= 12000, "rise_map: bad sizes");
    B200X_REQUIRE(n_freq == 1025, "rise_map: the cell index is frame * 1025 + bin (got n_freq=%d)", n_freq);
    dim3 grid(ceil_div(n_time, 256), n_freq);
    const double denom = static_cast<double>(n_masks) * keep_probability + 1e-8;
    rise_map_kernel<<<grid, 256, static_cast<size_t>(n_masks > 0 ? n_masks : 1) * sizeof(uint32_t), static_cast<cudaStream_t>(stream)>>>(
        d_pred, n_masks, seed, rise_threshold(keep_probability), denom, n_freq, n_time, d_map);
    B200X_CUDA_TRY(cudaGetLastError());
    return B200X_OK;
}

extern "C" int b200x_saliency_reduce(const int32_t* d_windows, const double* d_delta, int n_windows, int n_freq,
                                     int n_time, double* d_map, void* stream) {
    B200X_REQUIRE(n_freq > 0 && n_time > 0 && n_windows >= 0, "saliency: bad sizes");
    B200X_REQUIRE(n_windows <= 12000, "saliency: too many windows (%d) for the shared-memory index list", n_windows);
    dim3 grid(ceil_div(n_time, 128), ceil_div(n_freq, 16));
    const size_t smem = static_cast<size_t>(n_windows > 0 ? n_windows : 1) * sizeof(int);
    saliency_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(d_windows, d_delta, n_windows, n_freq, n_time, d_map);
    B200X_CUDA_TRY(cudaGetLastError());
    return B200X_OK;
}

extern "C" int b200x_band_map(const int32_t* d_band_rows, const double* d_delta, int n_bands, int n_freq, int n_time,
                              double* d_map, void* stream) {
    B200X_REQUIRE(n_freq > 0 && n_time > 0 && n_bands >= 0, "band_map: bad sizes");
    dim3 grid(ceil_div(n_time, 1024), n_freq);
    band_map_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_band_rows, d_delta, n_bands, n_freq, n_time, d_map);
    B200X_CUDA_TRY(cudaGetLastError());
    return B200X_OK;
}

extern "C" int b200x_rank(const double* d_values, int n, int mode, int32_t* d_order, void* stream) {
    B200X_REQUIRE(mode >= 0 && mode <= 3, "rank: bad mode %d", mode);
    if (n <= 0) return B200X_OK;
    rank_kernel<<<ceil_div(n, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(d_values, n, mode, d_order);
    B200X_CUDA_TRY(cudaGetLastError());
    return B200X_OK;
}
