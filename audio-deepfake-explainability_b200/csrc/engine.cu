// Engine: host-side orchestration of the perturbation sweep behind the C ABI (include/b200xai.h).
//
// One engine per GPU / process.  Per track: wave and explainer STFT stay resident in HBM.  A sweep of N perturbed
// copies is processed in chunks of `copies_per_chunk`; the perturbed spectrograms themselves are never materialised
// (the mask is generated in the iSTFT load stage).  Chunks should be LARGE (>= 100 copies, ~22 MB of workspace per
// copy at 120 s / 16 kHz): every kernel of the forward is a persistent / one-wave-per-SM launch whose prologue, tail
// and launch gap cost microseconds, so 16-copy chunks sized for L2 residency measured 25 % slower than one 228-copy
// chunk (tools/phase_times.py).  Everything is enqueued on one stream; no allocation in steady state.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "common.h"

using namespace b200x;

namespace {

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    int alloc(size_t n) {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = n;
        if (n == 0) return B200X_OK;
        cudaError_t e = cudaMalloc(&p, n);
        if (e != cudaSuccess) return set_error(B200X_ERR_CUDA, "cudaMalloc(%zu) failed: %s", n, cudaGetErrorString(e));
        return B200X_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct LayerW {
    DevBuf qkv_w, qkv_b, proj_w, proj_b, fc1_w, fc1_b, fc2_w, fc2_b, n1_g, n1_b, n2_g, n2_b;
};

inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

// pick the UMMA N tile that wastes the fewest padded columns (ties -> wider tile)
int pick_block_n(int n, int k = 1 << 30) {
    (void)k;
    const int cands[4] = {256, 208, 192, 128};
    int best = 128, best_waste = 1 << 30;
    for (int c : cands) {
        const int waste = ceil_div(n, c) * c - n;
        if (waste < best_waste) { best = c; best_waste = waste; }
    }
    return best;
}

}  // namespace

struct b200x_engine {
    b200x_model_config cfg;
    int device = 0;            // the CUDA device this engine lives on (made current by every entry point)
    int C = 0;                 // copies per chunk
    int64_t max_samples = 0;
    int T = 0, Tt = 0, Ts = 0; // tokens total / temporal / spectral
    int D = 0, Hp = 0;         // embed dim, padded MLP hidden
    bool finalized = false;
    cudaStream_t stream = nullptr;
    int64_t launches = 0;
    std::map<std::string, std::vector<float>> params;

    // packed weights
    DevBuf tok_t_w, tok_s_w, tok_t_b, tok_s_b, pe_t, pe_s, np_t_g, np_t_b, np_s_g, np_s_b;
    std::vector<LayerW> layers;
    DevBuf fn_g, fn_b, cls_w;
    float cls_b = 0.f;

    // track state
    int64_t L = 0;             // samples of the current track
    int n_time = 0;
    static constexpr int n_freq = 1025;
    static constexpr int s_stride = 1028;
    DevBuf wave, S;
    double ref_rms = 0.0;      // sqrt(mean(wave^2) + 1e-8)
    // baseline of the current track for the sparse occlusion path: dB mel of istft(S) and its prefix / suffix maxima
    DevBuf db_base, base_pre, base_suf, ranges;
    bool baseline_valid = false;

    // workspace
    int64_t y_stride = 0;
    int max_frames = 0, n_cta_max = 0;
    DevBuf y, db, cta_max, partial, floor_v, img_t, img_f, x, h, qkv, att, hid, head_part, prob, logit, sumsq;
    DevBuf windows, gains, masks, stems, delta, order, map;
    DevBuf S_multi, ref_arr;   // batch-of-tracks FBP: the tracks' spectrograms back to back, per-copy match_rms reference levels
    const double* ref_arr_cur = nullptr;
    int32_t base_range_host[2] = {0, 0};   // frame range of a baseline copy appended to a sparse chunk (async upload source)
    bool alternate = true;     // flip the traversal direction between consecutive kernels of the forward (L2 reuse)   // non-null while a multi-track chunk is in flight (forward_chunk_body -> mel)
    int last_copies = 0;
    float* trace = nullptr;

    // mel-domain explainer variant (spec_type: mel): Slaney filterbank operators, the track's power mel spectrogram, the
    // baseline NNLS magnitude and the Griffin-Lim workspace of one chunk of copies (allocated on first use)
    int mel_n = 0;             // n_mels of the uploaded basis (0 = none)
    float mel_step = 0.f;
    DevBuf mel_basis, mel_bin_range, mel_bin_first, mel_bin_w, mel_pinv_t, mel_track, mel_mag_base, mel_frames;
    DevBuf gl_mag, gl_c, gl_r0, gl_r1;
    int gl_copies = 0;         // copies the Griffin-Lim workspace holds
    int mel_track_frames = 0;  // n_time the cached mel_track / mel_mag_base belong to (0 = stale)
    int mel_mag_iter = -1;     // nnls_iter the cached baseline magnitude was solved with

    // CUDA graphs of the classifier forward, one per chunk shape: the ~90 launches of a chunk are replayed with one
    // cudaGraphLaunch (inter-kernel gaps shrink, no host work per kernel).  state 0 = unseen (run eagerly once: lazy
    // one-time initialisation must not happen inside a capture), 1 = warmed (capture on the next use), 2 = ready.
    struct ChunkGraph { int state = 0; cudaGraphExec_t exec = nullptr; int64_t launches = 0; uint64_t last_use = 0; };
    std::map<std::tuple<int, int64_t, int, int>, ChunkGraph> graphs;
    uint64_t graph_clock = 0;
    static constexpr size_t max_graphs = 16;   // least-recently-used shapes beyond this are destroyed (tracks of many lengths)
    bool use_graphs = true;
    bool spectral_mmajor = false;   // n_mels == 128, f_clip == 1: no transposed image (b200x_gemm_tokens_mmajor)
    bool fuse_ln = true;       // LayerNorm as a tail of the residual GEMM before it (gemm_resid_ln) instead of a pass of its own
    DevBuf prob_chunk, logit_chunk, ranges_chunk;      // fixed addresses baked into the graphs

    // optional per-kernel-class CUDA-event timing (bench roofline breakdown)
    bool timing = false;
    std::vector<cudaEvent_t> ev_pool;
    std::vector<int> ev_class;     // class of the i-th (start, stop) pair
    size_t ev_used = 0;
};

enum KernelClass { KC_ISTFT = 0, KC_MEL = 1, KC_RESIZE = 2, KC_GEMM = 3, KC_ATTN = 4, KC_LN = 5, KC_HEAD = 6, KC_OTHER = 7, KC_COUNT = 8 };

namespace {

// RAII-free bracket: records a (start, stop) event pair around a launch when timing is enabled
struct Timed {
    b200x_engine* e;
    size_t slot = 0;
    bool on = false;
    Timed(b200x_engine* eng, int cls) : e(eng) {
        if (!e->timing) return;
        if (e->ev_used + 2 > e->ev_pool.size()) {
            for (int i = 0; i < 2; ++i) { cudaEvent_t ev; cudaEventCreate(&ev); e->ev_pool.push_back(ev); }
        }
        slot = e->ev_used;
        e->ev_used += 2;
        e->ev_class.push_back(cls);
        cudaEventRecord(e->ev_pool[slot], e->stream);
        on = true;
    }
    ~Timed() { if (on) cudaEventRecord(e->ev_pool[slot + 1], e->stream); }
};
#define TIMED(cls, call) do { Timed _t(e, cls); B200X_TRY(call); } while (0)

int upload(DevBuf& b, const void* host, size_t bytes) {
    B200X_TRY(b.alloc(bytes));
    B200X_CUDA_TRY(cudaMemcpy(b.p, host, bytes, cudaMemcpyHostToDevice));
    return B200X_OK;
}

int upload_bf16(DevBuf& b, const std::vector<float>& v) {
    std::vector<__nv_bfloat16> t(v.size());
    for (size_t i = 0; i < v.size(); ++i) t[i] = __float2bfloat16_rn(v[i]);
    return upload(b, t.data(), t.size() * sizeof(__nv_bfloat16));
}

int get_param(b200x_engine* e, const std::string& name, size_t numel, const std::vector<float>** out) {
    auto it = e->params.find(name);
    if (it == e->params.end()) return set_error(B200X_ERR_STATE, "missing parameter '%s'", name.c_str());
    if (it->second.size() != numel)
        return set_error(B200X_ERR_INVALID, "parameter '%s' has %zu elements, expected %zu", name.c_str(), it->second.size(), numel);
    *out = &it->second;
    return B200X_OK;
}

int ensure_grow(DevBuf& b, size_t bytes) {
    if (b.bytes >= bytes && b.p) return B200X_OK;
    return b.alloc(bytes);
}

// The SpecTTTra forward over `copies` waves already sitting in e->y (rows of y_stride floats, n_samples valid).
int forward_chunk_body(b200x_engine* e, int copies, int64_t n_samples, const double* d_sumsq, int64_t rms_count, float* d_prob,
                       float* d_logit, const int32_t* d_ranges, int max_range) {
    const b200x_model_config& c = e->cfg;
    cudaStream_t s = e->stream;
    const int n_frames = 1 + static_cast<int>(n_samples / c.hop_length);
    const int span = d_ranges ? std::min(n_frames, std::max(1, max_range)) : n_frames;
    const int n_cta = ceil_div(span, b200x_mel_frames_per_cta());
    const int D = e->D, T = e->T, M = copies * T;
    e->last_copies = copies;
    TIMED(KC_MEL, b200x_mel_db_ref(e->y.as<float>(), e->y_stride, n_samples, copies, c.sample_rate, c.n_mels, c.f_min, c.f_max, c.amin,
                               d_sumsq, e->ref_rms, e->ref_arr_cur, rms_count, e->db.as<float>(), e->max_frames,
                               e->cta_max.as<float>(), d_ranges, max_range, s));
    TIMED(KC_RESIZE, b200x_mel_normalize_resize(e->db.as<float>(), e->max_frames, e->cta_max.as<float>(), n_cta, copies, n_frames,
                                         c.n_mels, static_cast<float>(c.top_db), c.std_unbiased, c.norm_eps, c.input_temp_dim,
                                         d_ranges ? e->db_base.as<float>() : nullptr, e->base_pre.as<float>(),
                                         e->base_suf.as<float>(), d_ranges, e->partial.p, e->floor_v.as<float>(), e->img_t.p,
                                         e->spectral_mmajor ? nullptr : e->img_f.p, c.input_temp_dim, s));
    e->launches += 3;
    // tokenizers: temporal rows = t_clip consecutive time steps x n_mels; spectral rows = one mel row over time
    const int Kt = c.t_clip * c.input_spec_dim;
    TIMED(KC_GEMM, b200x_gemm_bf16(e->img_t.p, Kt, e->tok_t_w.p, Kt, copies * e->Tt, D, Kt, pick_block_n(D), e->x.p, D,
                              B200X_GEMM_OUT_F32_TOKEN, e->tok_t_b.as<float>(), 1, nullptr, e->pe_t.as<float>(), e->Tt, T, 0, 0, s));
    if (e->spectral_mmajor) {
        // 128 mel rows = one 128-row MMA tile per copy: the spectral tokenizer reads the [time][mel] image through an M-major
        // operand descriptor; the transposed [mel][time] copy (img_f) is never written
        TIMED(KC_GEMM, b200x_gemm_tokens_mmajor(e->img_t.p, copies, c.input_temp_dim, e->tok_s_w.p, c.input_temp_dim, D, e->x.as<float>(), D,
                                           e->tok_s_b.as<float>(), 1, e->pe_s.as<float>(), T, e->Tt, s));
    } else {
        TIMED(KC_GEMM, b200x_gemm_bf16(e->img_f.p, c.input_temp_dim, e->tok_s_w.p, c.input_temp_dim, copies * e->Ts, D,
                                  c.input_temp_dim, pick_block_n(D), e->x.p, D, B200X_GEMM_OUT_F32_TOKEN, e->tok_s_b.as<float>(), 1,
                                  nullptr, e->pe_s.as<float>(), e->Ts, T, e->Tt, 0, s));
    }
    e->launches += 2;
    if (c.pre_norm) {
        TIMED(KC_LN, b200x_layernorm(e->x.as<float>(), M, D, e->np_t_g.as<float>(), e->np_t_b.as<float>(), e->np_s_g.as<float>(),
                                  e->np_s_b.as<float>(), T, e->Tt, c.tokenizer_ln_eps, nullptr, e->x.as<float>(), 0, s));
        e->launches += 1;
    }
    const size_t xbytes = static_cast<size_t>(M) * D * sizeof(float);
    if (e->trace) B200X_CUDA_TRY(cudaMemcpyAsync(e->trace, e->x.p, xbytes, cudaMemcpyDeviceToDevice, s));
    // Alternating traversal: consecutive kernels of the forward walk their rows / tiles / (copy, head) blocks in opposite
    // directions, so each one starts on the part of its input that its producer wrote last and that is still in L2 (the
    // 229-copy activations are 0.25-1.1 GB per tensor).  The direction is a launch ARGUMENT (no process-wide state).
    int rev = 1;                          // the first LayerNorm walks forward (rev flips to 0 before its launch)
    auto dir = [&]() { if (e->alternate) rev ^= 1; else rev = 0; return rev; };
    // With the LayerNorm tail (default) every residual GEMM leaves h = LayerNorm(x) for the projection that follows it
    // (b200x_gemm_resid_ln_bf16: proj -> LN2 -> fc1 with the load-add-store epilogue, fc2 -> next block's LN1 -> QKV with the
    // reduce-add epilogue); only the first block's LN1 is a pass of its own.
    const bool fuse = e->fuse_ln && D % 128 == 0 && D <= 384;
    for (int l = 0; l < c.num_layers; ++l) {
        LayerW& w = e->layers[l];
        const bool last = l + 1 == c.num_layers;
        if (!fuse || l == 0) {
            TIMED(KC_LN, b200x_layernorm(e->x.as<float>(), M, D, w.n1_g.as<float>(), w.n1_b.as<float>(), nullptr, nullptr, 0, 0,
                                      c.block_ln_eps, e->h.p, nullptr, dir(), s));
            e->launches += 1;
        }
        TIMED(KC_GEMM, b200x_gemm_bf16(e->h.p, D, w.qkv_w.p, D, M, 3 * D, D, pick_block_n(3 * D, D), e->qkv.p, 3 * D, B200X_GEMM_OUT_BF16,
                                  c.qkv_bias ? w.qkv_b.as<float>() : nullptr, 0, nullptr, nullptr, 0, 0, 0, dir(), s));
        TIMED(KC_ATTN, b200x_attention(e->qkv.p, e->att.p, copies, T, c.num_heads, D / c.num_heads, dir(), s));
        if (fuse) {
            TIMED(KC_GEMM, b200x_gemm_resid_ln_bf16(e->att.p, D, w.proj_w.p, D, M, D, D, e->x.as<float>(), D, w.proj_b.as<float>(),
                                               w.n2_g.as<float>(), w.n2_b.as<float>(), c.block_ln_eps, e->h.p, D, dir(), s));
        } else {
            TIMED(KC_GEMM, b200x_gemm_bf16(e->att.p, D, w.proj_w.p, D, M, D, D, pick_block_n(D, D), e->x.p, D, B200X_GEMM_OUT_F32_RESID,
                                      w.proj_b.as<float>(), 0, e->x.as<float>(), nullptr, 0, 0, 0, dir(), s));
            TIMED(KC_LN, b200x_layernorm(e->x.as<float>(), M, D, w.n2_g.as<float>(), w.n2_b.as<float>(), nullptr, nullptr, 0, 0,
                                      c.block_ln_eps, e->h.p, nullptr, dir(), s));
            e->launches += 1;
        }
        TIMED(KC_GEMM, b200x_gemm_bf16(e->h.p, D, w.fc1_w.p, D, M, e->Hp, D, pick_block_n(e->Hp, D), e->hid.p, e->Hp, B200X_GEMM_OUT_BF16,
                                  w.fc1_b.as<float>(), 1, nullptr, nullptr, 0, 0, 0, dir(), s));
        if (fuse && !last) {
            LayerW& wn = e->layers[l + 1];
            TIMED(KC_GEMM, b200x_gemm_resid_ln_bf16(e->hid.p, e->Hp, w.fc2_w.p, e->Hp, M, D, e->Hp, e->x.as<float>(), D, w.fc2_b.as<float>(),
                                               wn.n1_g.as<float>(), wn.n1_b.as<float>(), c.block_ln_eps, e->h.p, D, dir(), s));
        } else {
            TIMED(KC_GEMM, b200x_gemm_bf16(e->hid.p, e->Hp, w.fc2_w.p, e->Hp, M, D, e->Hp, pick_block_n(D), e->x.p, D,
                                      B200X_GEMM_OUT_F32_RESID, w.fc2_b.as<float>(), 0, e->x.as<float>(), nullptr, 0, 0, 0, dir(), s));
        }
        e->launches += 5;
        if (e->trace)
            B200X_CUDA_TRY(cudaMemcpyAsync(e->trace + static_cast<size_t>(l + 1) * M * D, e->x.p, xbytes, cudaMemcpyDeviceToDevice, s));
    }
    TIMED(KC_HEAD, b200x_head(e->x.as<float>(), copies, T, D, e->fn_g.as<float>(), e->fn_b.as<float>(), c.block_ln_eps, c.final_norm,
                         e->cls_w.as<float>(), e->cls_b, e->head_part.as<float>(), d_logit, d_prob, s));
    e->launches += 2;
    return B200X_OK;
}

// forward_chunk = forward_chunk_body, replayed from a CUDA graph once the chunk shape has been seen twice
int forward_chunk(b200x_engine* e, int copies, int64_t n_samples, const double* d_sumsq, int64_t rms_count, float* d_prob,
                  float* d_logit, const int32_t* d_ranges = nullptr, int max_range = 0) {
    // the loudness-normalised FBP path bakes a per-track scalar (ref_rms) into its launches: it stays eager
    const bool graphable = e->use_graphs && !e->timing && e->trace == nullptr && d_sumsq == nullptr;
    if (!graphable) return forward_chunk_body(e, copies, n_samples, d_sumsq, rms_count, d_prob, d_logit, d_ranges, max_range);
    cudaStream_t s = e->stream;
    if (d_ranges != nullptr)
        B200X_CUDA_TRY(cudaMemcpyAsync(e->ranges_chunk.p, d_ranges, static_cast<size_t>(copies) * 2 * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
    const int32_t* g_ranges = d_ranges ? e->ranges_chunk.as<int32_t>() : nullptr;
    float* g_prob = e->prob_chunk.as<float>();
    float* g_logit = e->logit_chunk.as<float>();
    const auto key = std::make_tuple(copies, n_samples, d_ranges ? 1 : 0, max_range);
    if (e->graphs.find(key) == e->graphs.end() && e->graphs.size() >= b200x_engine::max_graphs) {
        auto victim = e->graphs.begin();                  // evict the least recently used shape
        for (auto it = e->graphs.begin(); it != e->graphs.end(); ++it)
            if (it->second.last_use < victim->second.last_use) victim = it;
        if (victim->second.exec) {
            B200X_CUDA_TRY(cudaStreamSynchronize(s));     // the victim may still be executing
            cudaGraphExecDestroy(victim->second.exec);
        }
        e->graphs.erase(victim);
    }
    b200x_engine::ChunkGraph& g = e->graphs[key];
    g.last_use = ++e->graph_clock;
    if (g.state == 0) {
        B200X_TRY(forward_chunk_body(e, copies, n_samples, nullptr, 0, g_prob, g_logit, g_ranges, max_range));
        g.state = 1;
    } else {
        if (g.state == 1) {
            const int64_t before = e->launches;
            B200X_CUDA_TRY(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
            const int rc = forward_chunk_body(e, copies, n_samples, nullptr, 0, g_prob, g_logit, g_ranges, max_range);
            cudaGraph_t graph = nullptr;
            const cudaError_t ce = cudaStreamEndCapture(s, &graph);
            g.launches = e->launches - before;
            e->launches = before;
            if (rc != B200X_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
            if (ce != cudaSuccess) return set_error(B200X_ERR_CUDA, "graph capture of the forward pass failed: %s", cudaGetErrorString(ce));
            const cudaError_t ie = cudaGraphInstantiate(&g.exec, graph, 0);
            cudaGraphDestroy(graph);
            if (ie != cudaSuccess) return set_error(B200X_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(ie));
            g.state = 2;
        }
        B200X_CUDA_TRY(cudaGraphLaunch(g.exec, s));
        e->launches += g.launches;
    }
    B200X_CUDA_TRY(cudaMemcpyAsync(d_prob, g_prob, static_cast<size_t>(copies) * sizeof(float), cudaMemcpyDeviceToDevice, s));
    if (d_logit != nullptr)
        B200X_CUDA_TRY(cudaMemcpyAsync(d_logit, g_logit, static_cast<size_t>(copies) * sizeof(float), cudaMemcpyDeviceToDevice, s));
    return B200X_OK;
}

// reference RMS for match_rms: sqrt(mean(sig^2) + 1e-8) in float64 on the host, like the reference (dsp_band_ops.py:228-233)
int ensure_ref_rms(b200x_engine* e) {
    if (e->ref_rms >= 0.0) return B200X_OK;
    B200X_TRY(ensure_grow(e->ref_arr, sizeof(double)));
    B200X_TRY(b200x_wave_rms(e->wave.as<float>(), e->L, e->L, 1, 1, e->ref_arr.as<double>(), e->stream));
    e->launches += 1;
    double r = 0.0;
    B200X_CUDA_TRY(cudaMemcpyAsync(&r, e->ref_arr.p, sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    B200X_CUDA_TRY(cudaStreamSynchronize(e->stream));
    e->ref_rms = r;
    return B200X_OK;
}

int check_ready(b200x_engine* e, bool need_track) {
    if (e == nullptr) return set_error(B200X_ERR_INVALID, "engine is NULL");
    B200X_CUDA_TRY(cudaSetDevice(e->device));         // several engines (one per GPU) may share a host thread
    if (!e->finalized) return set_error(B200X_ERR_STATE, "engine weights not finalized");
    if (need_track && e->L == 0) return set_error(B200X_ERR_STATE, "no track loaded (call b200x_engine_set_track)");
    return B200X_OK;
}

int ensure_prob(b200x_engine* e, int n) {
    B200X_TRY(ensure_grow(e->prob, static_cast<size_t>(std::max(n, 1)) * sizeof(float)));
    B200X_TRY(ensure_grow(e->logit, static_cast<size_t>(std::max(n, 1)) * sizeof(float)));
    return B200X_OK;
}

}  // namespace

extern "C" int b200x_engine_create(const b200x_model_config* cfg, int copies_per_chunk, int64_t max_samples,
                                   b200x_engine** out) {
    B200X_REQUIRE(cfg != nullptr && out != nullptr, "engine_create: NULL argument");
    B200X_REQUIRE(cfg->n_fft == 2048 && cfg->hop_length == 512, "engine: only n_fft=2048 / hop=512 kernels are built");
    B200X_REQUIRE(cfg->embed_dim % 128 == 0 && cfg->embed_dim / cfg->num_heads == 64, "engine: embed_dim must be a multiple of 128 with head_dim 64");
    B200X_REQUIRE(cfg->f_clip == 1, "engine: f_clip=%d unsupported (alpha variant uses 1)", cfg->f_clip);
    B200X_REQUIRE(cfg->input_spec_dim == cfg->n_mels, "engine: input_spec_dim must equal n_mels (no resize along mel axis)");
    B200X_REQUIRE((cfg->t_clip * cfg->input_spec_dim) % 8 == 0 && cfg->input_temp_dim % 8 == 0, "engine: tokenizer K not 16-byte aligned");
    B200X_REQUIRE(copies_per_chunk >= 1 && copies_per_chunk <= 1024, "engine: copies_per_chunk out of range");
    B200X_REQUIRE(max_samples >= 4096, "engine: max_samples too small");
    int dev = 0, major = 0;
    B200X_CUDA_TRY(cudaGetDevice(&dev));
    B200X_CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    B200X_REQUIRE(major == 10, "engine: this library contains sm_100a code only (device has compute capability %d.x)", major);
    b200x_engine* e = new b200x_engine();
    e->device = dev;
    e->cfg = *cfg;
    e->C = copies_per_chunk;
    e->max_samples = max_samples;
    e->Tt = (cfg->input_temp_dim - cfg->t_clip) / cfg->t_clip + 1;
    e->Ts = (cfg->input_spec_dim - cfg->f_clip) / cfg->f_clip + 1;
    e->T = e->Tt + e->Ts;
    e->D = cfg->embed_dim;
    e->Hp = round_up(cfg->mlp_hidden, 16);
    if (e->T % 16 != 0) { delete e; return set_error(B200X_ERR_INVALID, "engine: token count %d must be a multiple of 16", e->T); }
    cudaError_t ce = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking);
    if (ce != cudaSuccess) { delete e; return set_error(B200X_ERR_CUDA, "stream create failed: %s", cudaGetErrorString(ce)); }
    const int C = e->C, D = e->D, T = e->T;
    const size_t M = static_cast<size_t>(C) * T;
    e->y_stride = (max_samples + 7) / 8 * 8;
    e->max_frames = 1 + static_cast<int>(max_samples / cfg->hop_length);
    e->n_cta_max = ceil_div(e->max_frames, b200x_mel_frames_per_cta());
    int st = B200X_OK;
    auto A = [&](DevBuf& b, size_t bytes) { if (st == B200X_OK) st = b.alloc(bytes); };
    A(e->y, static_cast<size_t>(C) * e->y_stride * sizeof(float));
    A(e->db, static_cast<size_t>(C) * e->max_frames * cfg->n_mels * sizeof(float));
    A(e->cta_max, static_cast<size_t>(C) * e->n_cta_max * sizeof(float));
    A(e->partial, static_cast<size_t>(C) * 32 * 2 * sizeof(double));
    A(e->floor_v, static_cast<size_t>(C) * sizeof(float));
    A(e->img_t, static_cast<size_t>(C) * cfg->input_temp_dim * cfg->n_mels * 2);
    e->spectral_mmajor = cfg->n_mels == 128 && cfg->input_spec_dim == 128 && cfg->f_clip == 1 && e->Ts == 128;
    if (!e->spectral_mmajor) A(e->img_f, static_cast<size_t>(C) * cfg->n_mels * cfg->input_temp_dim * 2);
    A(e->x, M * D * sizeof(float));
    A(e->h, M * D * 2);
    A(e->qkv, M * 3 * D * 2);
    A(e->att, M * D * 2);
    A(e->hid, M * e->Hp * 2);
    A(e->head_part, static_cast<size_t>(C) * b200x_head_slices() * sizeof(float));
    A(e->sumsq, static_cast<size_t>(C) * sizeof(double));
    A(e->wave, static_cast<size_t>(e->y_stride) * sizeof(float));
    A(e->S, static_cast<size_t>(e->max_frames) * b200x_engine::s_stride * 2 * sizeof(float));
    A(e->db_base, static_cast<size_t>(e->max_frames) * cfg->n_mels * sizeof(float));
    A(e->base_pre, static_cast<size_t>(e->max_frames + 1) * sizeof(float));
    A(e->base_suf, static_cast<size_t>(e->max_frames + 1) * sizeof(float));
    A(e->prob_chunk, static_cast<size_t>(C) * sizeof(float));
    A(e->logit_chunk, static_cast<size_t>(C) * sizeof(float));
    A(e->ranges_chunk, static_cast<size_t>(C) * 2 * sizeof(int32_t));
    if (st != B200X_OK) { b200x_engine_destroy(e); return st; }
    cudaMemset(e->y.p, 0, e->y.bytes);
    *out = e;
    return B200X_OK;
}

extern "C" void b200x_engine_destroy(b200x_engine* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    for (auto& kv : e->graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    e->graphs.clear();
    DevBuf* bufs[] = {&e->tok_t_w, &e->tok_s_w, &e->tok_t_b, &e->tok_s_b, &e->pe_t, &e->pe_s, &e->np_t_g, &e->np_t_b, &e->np_s_g,
                      &e->np_s_b, &e->fn_g, &e->fn_b, &e->cls_w, &e->wave, &e->S, &e->y, &e->db, &e->cta_max, &e->partial,
                      &e->floor_v, &e->img_t, &e->img_f, &e->x, &e->h, &e->qkv, &e->att, &e->hid, &e->head_part, &e->prob,
                      &e->logit, &e->sumsq, &e->db_base, &e->base_pre, &e->base_suf, &e->ranges, &e->prob_chunk, &e->logit_chunk, &e->ranges_chunk, &e->windows, &e->gains, &e->masks, &e->stems, &e->delta, &e->order, &e->map,
                      &e->S_multi, &e->ref_arr, &e->mel_basis, &e->mel_bin_range, &e->mel_bin_first, &e->mel_bin_w, &e->mel_pinv_t, &e->mel_track,
                      &e->mel_mag_base, &e->mel_frames, &e->gl_mag, &e->gl_c, &e->gl_r0, &e->gl_r1};
    for (DevBuf* b : bufs) b->release();
    for (LayerW& w : e->layers) {
        DevBuf* lb[] = {&w.qkv_w, &w.qkv_b, &w.proj_w, &w.proj_b, &w.fc1_w, &w.fc1_b, &w.fc2_w, &w.fc2_b, &w.n1_g, &w.n1_b, &w.n2_g, &w.n2_b};
        for (DevBuf* b : lb) b->release();
    }
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

extern "C" int b200x_engine_set_param(b200x_engine* e, const char* name, const float* data, int64_t numel) {
    B200X_REQUIRE(e && name && data && numel > 0, "set_param: bad argument");
    e->params[name] = std::vector<float>(data, data + numel);
    e->finalized = false;
    return B200X_OK;
}

extern "C" int b200x_engine_finalize(b200x_engine* e) {
    B200X_REQUIRE(e != nullptr, "finalize: engine is NULL");
    B200X_CUDA_TRY(cudaSetDevice(e->device));
    const b200x_model_config& c = e->cfg;
    const int D = e->D, F = c.input_spec_dim, Tm = c.input_temp_dim, H = c.mlp_hidden, Hp = e->Hp;
    const std::vector<float>* p = nullptr;
    const std::string tk = "encoder.st_tokenizer.";
    // temporal conv weight [D][F][t_clip] -> [D][t_clip * F], K index = k * F + f (matches img_t rows)
    B200X_TRY(get_param(e, tk + "temporal_tokenizer.conv1d.weight", static_cast<size_t>(D) * F * c.t_clip, &p));
    {
        std::vector<float> w(static_cast<size_t>(D) * c.t_clip * F);
        for (int o = 0; o < D; ++o)
            for (int f = 0; f < F; ++f)
                for (int k = 0; k < c.t_clip; ++k)
                    w[(static_cast<size_t>(o) * c.t_clip + k) * F + f] = (*p)[(static_cast<size_t>(o) * F + f) * c.t_clip + k];
        B200X_TRY(upload_bf16(e->tok_t_w, w));
    }
    B200X_TRY(get_param(e, tk + "spectral_tokenizer.conv1d.weight", static_cast<size_t>(D) * Tm * c.f_clip, &p));
    B200X_TRY(upload_bf16(e->tok_s_w, *p));
    std::vector<float> zeros_d(D, 0.f), ones_d(D, 1.f);
    if (!c.pre_norm) {
        B200X_TRY(get_param(e, tk + "temporal_tokenizer.conv1d.bias", D, &p));
        B200X_TRY(upload(e->tok_t_b, p->data(), D * sizeof(float)));
        B200X_TRY(get_param(e, tk + "spectral_tokenizer.conv1d.bias", D, &p));
        B200X_TRY(upload(e->tok_s_b, p->data(), D * sizeof(float)));
    } else {
        B200X_TRY(upload(e->tok_t_b, zeros_d.data(), D * sizeof(float)));
        B200X_TRY(upload(e->tok_s_b, zeros_d.data(), D * sizeof(float)));
        B200X_TRY(get_param(e, tk + "temporal_tokenizer.norm_pre.weight", D, &p)); B200X_TRY(upload(e->np_t_g, p->data(), D * 4));
        B200X_TRY(get_param(e, tk + "temporal_tokenizer.norm_pre.bias", D, &p));   B200X_TRY(upload(e->np_t_b, p->data(), D * 4));
        B200X_TRY(get_param(e, tk + "spectral_tokenizer.norm_pre.weight", D, &p)); B200X_TRY(upload(e->np_s_g, p->data(), D * 4));
        B200X_TRY(get_param(e, tk + "spectral_tokenizer.norm_pre.bias", D, &p));   B200X_TRY(upload(e->np_s_b, p->data(), D * 4));
    }
    if (c.pe_learnable) {
        B200X_TRY(get_param(e, tk + "temporal_tokenizer.pos_encoder.pe", static_cast<size_t>(e->Tt) * D, &p));
        B200X_TRY(upload(e->pe_t, p->data(), p->size() * 4));
        B200X_TRY(get_param(e, tk + "spectral_tokenizer.pos_encoder.pe", static_cast<size_t>(e->Ts) * D, &p));
        B200X_TRY(upload(e->pe_s, p->data(), p->size() * 4));
    } else {                                   // sinusoidal encoding, computed here in float32 like torch
        for (int which = 0; which < 2; ++which) {
            const int n = which == 0 ? e->Tt : e->Ts;
            std::vector<float> pe(static_cast<size_t>(n) * D, 0.f);
            for (int pos = 0; pos < n; ++pos)
                for (int i = 0; i < D; i += 2) {
                    const float div = std::exp(static_cast<float>(i) * static_cast<float>(-std::log(10000.0) / D));
                    pe[static_cast<size_t>(pos) * D + i] = std::sin(pos * div);
                    if (i + 1 < D) pe[static_cast<size_t>(pos) * D + i + 1] = std::cos(pos * div);
                }
            B200X_TRY(upload(which == 0 ? e->pe_t : e->pe_s, pe.data(), pe.size() * 4));
        }
    }
    e->layers.clear();
    e->layers.resize(c.num_layers);
    for (int l = 0; l < c.num_layers; ++l) {
        LayerW& w = e->layers[l];
        const std::string b = "encoder.transformer.blocks." + std::to_string(l) + ".";
        B200X_TRY(get_param(e, b + "norm1.weight", D, &p)); B200X_TRY(upload(w.n1_g, p->data(), D * 4));
        B200X_TRY(get_param(e, b + "norm1.bias", D, &p));   B200X_TRY(upload(w.n1_b, p->data(), D * 4));
        B200X_TRY(get_param(e, b + "norm2.weight", D, &p)); B200X_TRY(upload(w.n2_g, p->data(), D * 4));
        B200X_TRY(get_param(e, b + "norm2.bias", D, &p));   B200X_TRY(upload(w.n2_b, p->data(), D * 4));
        B200X_TRY(get_param(e, b + "attn.qkv.weight", static_cast<size_t>(3) * D * D, &p)); B200X_TRY(upload_bf16(w.qkv_w, *p));
        if (c.qkv_bias) { B200X_TRY(get_param(e, b + "attn.qkv.bias", 3 * D, &p)); B200X_TRY(upload(w.qkv_b, p->data(), 3 * D * 4)); }
        B200X_TRY(get_param(e, b + "attn.proj.weight", static_cast<size_t>(D) * D, &p)); B200X_TRY(upload_bf16(w.proj_w, *p));
        B200X_TRY(get_param(e, b + "attn.proj.bias", D, &p)); B200X_TRY(upload(w.proj_b, p->data(), D * 4));
        // MLP hidden padded to a multiple of 16 with zero rows / columns (GELU(0) = 0 keeps the padding inert)
        B200X_TRY(get_param(e, b + "mlp.fc1.weight", static_cast<size_t>(H) * D, &p));
        { std::vector<float> t(static_cast<size_t>(Hp) * D, 0.f); std::copy(p->begin(), p->end(), t.begin()); B200X_TRY(upload_bf16(w.fc1_w, t)); }
        B200X_TRY(get_param(e, b + "mlp.fc1.bias", H, &p));
        { std::vector<float> t(Hp, 0.f); std::copy(p->begin(), p->end(), t.begin()); B200X_TRY(upload(w.fc1_b, t.data(), Hp * 4)); }
        B200X_TRY(get_param(e, b + "mlp.fc2.weight", static_cast<size_t>(D) * H, &p));
        { std::vector<float> t(static_cast<size_t>(D) * Hp, 0.f);
          for (int o = 0; o < D; ++o) std::copy(p->begin() + static_cast<size_t>(o) * H, p->begin() + static_cast<size_t>(o + 1) * H, t.begin() + static_cast<size_t>(o) * Hp);
          B200X_TRY(upload_bf16(w.fc2_w, t)); }
        B200X_TRY(get_param(e, b + "mlp.fc2.bias", D, &p)); B200X_TRY(upload(w.fc2_b, p->data(), D * 4));
    }
    if (c.final_norm) {
        B200X_TRY(get_param(e, "encoder.transformer.norm.weight", D, &p)); B200X_TRY(upload(e->fn_g, p->data(), D * 4));
        B200X_TRY(get_param(e, "encoder.transformer.norm.bias", D, &p));   B200X_TRY(upload(e->fn_b, p->data(), D * 4));
    } else {
        B200X_TRY(upload(e->fn_g, ones_d.data(), D * 4));
        B200X_TRY(upload(e->fn_b, zeros_d.data(), D * 4));
    }
    B200X_TRY(get_param(e, "classifier.weight", D, &p)); B200X_TRY(upload(e->cls_w, p->data(), D * 4));
    B200X_TRY(get_param(e, "classifier.bias", 1, &p));
    e->cls_b = (*p)[0];
    e->params.clear();
    for (auto& kv : e->graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    e->graphs.clear();                         // the weight buffers were re-allocated: captured pointers are stale
    e->finalized = true;
    return B200X_OK;
}

extern "C" int b200x_engine_predict(b200x_engine* e, const float* waves, int64_t n_samples, int count, int on_device,
                                    float* prob, float* logit) {
    B200X_TRY(check_ready(e, false));
    B200X_REQUIRE(waves && prob && count > 0, "predict: bad argument");
    B200X_REQUIRE(n_samples > 1024 && n_samples <= e->max_samples, "predict: n_samples=%lld outside (1024, %lld]", (long long)n_samples, (long long)e->max_samples);
    B200X_TRY(ensure_prob(e, count));
    const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    for (int c0 = 0; c0 < count; c0 += e->C) {
        const int n = std::min(e->C, count - c0);
        B200X_CUDA_TRY(cudaMemcpy2DAsync(e->y.p, e->y_stride * sizeof(float), waves + static_cast<size_t>(c0) * n_samples,
                                         n_samples * sizeof(float), n_samples * sizeof(float), n, kind, e->stream));
        B200X_TRY(forward_chunk(e, n, n_samples, nullptr, 0, e->prob.as<float>() + c0, e->logit.as<float>() + c0));
    }
    const cudaMemcpyKind back = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    B200X_CUDA_TRY(cudaMemcpyAsync(prob, e->prob.p, count * sizeof(float), back, e->stream));
    if (logit) B200X_CUDA_TRY(cudaMemcpyAsync(logit, e->logit.p, count * sizeof(float), back, e->stream));
    B200X_CUDA_TRY(cudaStreamSynchronize(e->stream));
    return B200X_OK;
}

// baseline prediction of the track loaded with set_track, from its device-resident samples (no second upload)
extern "C" int b200x_engine_predict_track(b200x_engine* e, float* prob, float* logit) {
    B200X_TRY(check_ready(e, true));
    B200X_REQUIRE(prob != nullptr, "predict_track: prob is NULL");
    B200X_TRY(ensure_prob(e, 1));
    B200X_CUDA_TRY(cudaMemcpyAsync(e->y.p, e->wave.p, e->L * sizeof(float), cudaMemcpyDeviceToDevice, e->stream));
    B200X_TRY(forward_chunk(e, 1, e->L, nullptr, 0, e->prob.as<float>(), e->logit.as<float>()));
    B200X_CUDA_TRY(cudaMemcpyAsync(prob, e->prob.p, sizeof(float), cudaMemcpyDeviceToHost, e->stream));
    if (logit) B200X_CUDA_TRY(cudaMemcpyAsync(logit, e->logit.p, sizeof(float), cudaMemcpyDeviceToHost, e->stream));
    B200X_CUDA_TRY(cudaStreamSynchronize(e->stream));
    return B200X_OK;
}

extern "C" int b200x_engine_set_track(b200x_engine* e, const float* wave, int64_t n_samples, int on_device) {
    B200X_TRY(check_ready(e, false));
    B200X_REQUIRE(wave != nullptr, "set_track: wave is NULL");
    B200X_REQUIRE(n_samples >= 2048 && n_samples <= e->max_samples, "set_track: n_samples=%lld outside [2048, %lld]", (long long)n_samples, (long long)e->max_samples);
    B200X_CUDA_TRY(cudaMemcpyAsync(e->wave.p, wave, n_samples * sizeof(float), on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, e->stream));
    e->L = n_samples;
    e->n_time = 1 + static_cast<int>(n_samples / e->cfg.hop_length);
    B200X_TRY(b200x_stft(e->wave.as<float>(), n_samples, e->cfg.n_fft, e->cfg.hop_length, 0, e->S.p, b200x_engine::s_stride, e->stream));
    e->launches += 1;
    e->baseline_valid = false;
    e->mel_track_frames = 0;
    e->ref_rms = -1.0;   // computed lazily (ensure_ref_rms) when a loudness-normalised FBP sweep asks for it
    // the tail of every y row beyond hop*(n_time-1) must read as zero padding (spectrogram_explainability.py:679-680)
    B200X_CUDA_TRY(cudaMemsetAsync(e->y.p, 0, e->y.bytes, e->stream));
    return B200X_OK;
}

extern "C" int b200x_engine_track_shape(b200x_engine* e, int32_t* n_freq, int32_t* n_time) {
    B200X_TRY(check_ready(e, true));
    if (n_freq) *n_freq = b200x_engine::n_freq;
    if (n_time) *n_time = e->n_time;
    return B200X_OK;
}

extern "C" int b200x_engine_get_spectrogram(b200x_engine* e, float* spec_host) {
    B200X_TRY(check_ready(e, true));
    B200X_REQUIRE(spec_host != nullptr, "get_spectrogram: NULL output");
    std::vector<float> tmp(static_cast<size_t>(e->n_time) * b200x_engine::s_stride * 2);
    B200X_CUDA_TRY(cudaMemcpyAsync(tmp.data(), e->S.p, tmp.size() * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
    B200X_CUDA_TRY(cudaStreamSynchronize(e->stream));
    const int F = b200x_engine::n_freq, Tn = e->n_time;
    for (int t = 0; t < Tn; ++t)
        for (int f = 0; f < F; ++f) {
            spec_host[(static_cast<size_t>(f) * Tn + t) * 2] = tmp[(static_cast<size_t>(t) * b200x_engine::s_stride + f) * 2];
            spec_host[(static_cast<size_t>(f) * Tn + t) * 2 + 1] = tmp[(static_cast<size_t>(t) * b200x_engine::s_stride + f) * 2 + 1];
        }
    return B200X_OK;
}

namespace {
// Baseline for the sparse occlusion path: y_base = istft(S) padded to len(y), its dB mel spectrogram and maxima.
int ensure_baseline(b200x_engine* e) {
    if (e->baseline_valid) return B200X_OK;
    const b200x_model_config& c = e->cfg;
    const int64_t out_len = static_cast<int64_t>(c.hop_length) * (e->n_time - 1);
    const int n_frames = 1 + static_cast<int>(e->L / c.hop_length);
    TIMED(KC_ISTFT, b200x_istft_masked(e->S.p, b200x_engine::s_stride, e->n_time, 1, B200X_MASK_NONE, nullptr, 0.f, nullptr,
                                 e->y.as<float>(), e->y_stride, nullptr, nullptr, 0, e->stream));
    if (e->L > out_len)
        B200X_CUDA_TRY(cudaMemsetAsync(e->y.as<float>() + out_len, 0, (e->L - out_len) * sizeof(float), e->stream));
    TIMED(KC_MEL, b200x_mel_db(e->y.as<float>(), e->y_stride, e->L, 1, c.sample_rate, c.n_mels, c.f_min, c.f_max, c.amin, nullptr, 0.0,
                         0, e->db_base.as<float>(), e->max_frames, e->cta_max.as<float>(), nullptr, 0, e->stream));
    TIMED(KC_OTHER, b200x_mel_base_maxima(e->db_base.as<float>(), n_frames, c.n_mels, e->base_pre.as<float>(), e->base_suf.as<float>(), e->stream));
    e->launches += 3;
    e->baseline_valid = true;
    return B200X_OK;
}

// shared body of the occlusion / FBP sweeps: perturb in the iSTFT load stage, classify, collect probabilities
// with_base: the unperturbed track itself is evaluated as one more copy of the LAST chunk (probability in d_prob_out[n]);
// on the sparse path that copy carries the full frame range, i.e. it is processed exactly like predict_track's single copy
int sweep(b200x_engine* e, int mode, int n, const int32_t* d_windows, float occ_value, const float* d_gains, bool rms,
          float* d_prob_out, int max_range = 0, bool with_base = false) {
    const int64_t out_len = static_cast<int64_t>(e->cfg.hop_length) * (e->n_time - 1);
    // occlusion: only classifier frames [t0-4, t1+4) differ from the unperturbed track (iSTFT linearity); everything
    // else is read from the per-track baseline.  Band gains change every frame, so FBP takes the dense path.
    const bool sparse = (mode == B200X_MASK_OCCLUDE) && max_range > 0 && max_range < e->n_time;
    const int32_t* d_ranges = nullptr;
    if (sparse) {
        B200X_TRY(ensure_baseline(e));
        B200X_TRY(ensure_grow(e->ranges, static_cast<size_t>(n + 1) * 2 * sizeof(int32_t)));
        const int n_frames_cls = 1 + static_cast<int>(e->L / e->cfg.hop_length);
        if (with_base) {
            e->base_range_host[0] = 0;
            e->base_range_host[1] = n_frames_cls;
            B200X_CUDA_TRY(cudaMemcpyAsync(e->ranges.as<int32_t>() + 2 * n, e->base_range_host, 2 * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
        }
        TIMED(KC_OTHER, b200x_frame_ranges(d_windows, n, n_frames_cls, e->ranges.as<int32_t>(), e->stream));
        e->launches += 1;
        d_ranges = e->ranges.as<int32_t>();
    }
    for (int c0 = 0; c0 < n; c0 += e->C) {
        const int m = std::min(e->C, n - c0);
        double* sumsq = nullptr;
        if (rms) {
            sumsq = e->sumsq.as<double>();
            B200X_CUDA_TRY(cudaMemsetAsync(sumsq, 0, m * sizeof(double), e->stream));
        }
        TIMED(KC_ISTFT, b200x_istft_masked(e->S.p, b200x_engine::s_stride, e->n_time, m, mode, d_windows ? d_windows + 4 * c0 : nullptr,
                                     occ_value, d_gains ? d_gains + static_cast<size_t>(c0) * b200x_engine::n_freq : nullptr,
                                     e->y.as<float>(), e->y_stride, sumsq, d_ranges ? d_ranges + 2 * c0 : nullptr, max_range, e->stream));
        e->launches += 1;
        // occlusion: the reference pads/trims y_occ to len(y); FBP feeds the iSTFT output as is (dsp_band_ops.py:580-586)
        const int64_t n_cls = (mode == B200X_MASK_BAND_GAIN) ? out_len : e->L;
        if (n_cls > out_len)   // zero padding of the tail (spectrogram_explainability.py:679-680)
            B200X_CUDA_TRY(cudaMemset2DAsync(e->y.as<float>() + out_len, e->y_stride * sizeof(float), 0,
                                             (n_cls - out_len) * sizeof(float), m, e->stream));
        const bool add_base = with_base && c0 + m == n && m < e->C && mode == B200X_MASK_OCCLUDE;
        if (add_base)    // row m of the chunk = the track itself (its tail beyond L is the zero padding every y row keeps)
            B200X_CUDA_TRY(cudaMemcpyAsync(e->y.as<float>() + static_cast<size_t>(m) * e->y_stride, e->wave.p, e->L * sizeof(float),
                                           cudaMemcpyDeviceToDevice, e->stream));
        const int chunk_range = (add_base && sparse) ? 1 + static_cast<int>(e->L / e->cfg.hop_length) : max_range;
        B200X_TRY(forward_chunk(e, m + (add_base ? 1 : 0), n_cls, sumsq, out_len, d_prob_out + c0, e->logit.as<float>() + c0,
                                d_ranges ? d_ranges + 2 * c0 : nullptr, chunk_range));
        if (with_base && c0 + m == n && !add_base) {     // no room in the last chunk: a chunk of its own, like predict_track
            B200X_CUDA_TRY(cudaMemcpyAsync(e->y.p, e->wave.p, e->L * sizeof(float), cudaMemcpyDeviceToDevice, e->stream));
            B200X_TRY(forward_chunk(e, 1, e->L, nullptr, 0, d_prob_out + n, e->logit.as<float>() + n));
        }
    }
    return B200X_OK;
}
}  // namespace

namespace {
int occlusion_sweep_impl(b200x_engine* e, const int32_t* windows, int n, float occlusion_value, int on_device, float* prob,
                         float* base_prob);
}

extern "C" int b200x_engine_occlusion_sweep(b200x_engine* e, const int32_t* windows, int n, float occlusion_value,
                                            int on_device, float* prob) {
    return occlusion_sweep_impl(e, windows, n, occlusion_value, on_device, prob, nullptr);
}

extern "C" int b200x_engine_occlusion_sweep_base(b200x_engine* e, const int32_t* windows, int n, float occlusion_value,
                                                 int on_device, float* prob, float* base_prob) {
    B200X_REQUIRE(base_prob != nullptr, "occlusion_sweep_base: base_prob is NULL");
    return occlusion_sweep_impl(e, windows, n, occlusion_value, on_device, prob, base_prob);
}

namespace {
int occlusion_sweep_impl(b200x_engine* e, const int32_t* windows, int n, float occlusion_value, int on_device, float* prob,
                         float* base_prob) {
    B200X_TRY(check_ready(e, true));
    if (n == 0 && base_prob == nullptr) return B200X_OK;
    if (n == 0) return b200x_engine_predict_track(e, base_prob, nullptr);
    B200X_REQUIRE(windows && prob && n > 0, "occlusion_sweep: bad argument");
    B200X_TRY(ensure_prob(e, n + 1));
    const int32_t* d_win = windows;
    std::vector<int32_t> host_copy;
    const int32_t* h_win = windows;
    if (on_device) {                       // the (tiny) window list is validated on the host in both cases
        host_copy.resize(static_cast<size_t>(n) * 4);
        B200X_CUDA_TRY(cudaMemcpyAsync(host_copy.data(), windows, static_cast<size_t>(n) * 16, cudaMemcpyDeviceToHost, e->stream));
        B200X_CUDA_TRY(cudaStreamSynchronize(e->stream));
        h_win = host_copy.data();
    }
    int max_range = 0;
    for (int i = 0; i < n; ++i) {
        const int32_t* w = h_win + 4 * i;
        B200X_REQUIRE(w[0] >= 0 && w[0] <= w[1] && w[1] <= e->n_time && w[2] >= 0 && w[2] <= w[3] && w[3] <= b200x_engine::n_freq,
                      "occlusion_sweep: window %d = (%d,%d,%d,%d) outside the %dx%d spectrogram", i, w[0], w[1], w[2], w[3],
                      b200x_engine::n_freq, e->n_time);
        max_range = std::max(max_range, w[1] - w[0] + 8);
    }
    if (!on_device) {
        B200X_TRY(ensure_grow(e->windows, static_cast<size_t>(n) * 4 * sizeof(int32_t)));
        B200X_CUDA_TRY(cudaMemcpyAsync(e->windows.p, windows, static_cast<size_t>(n) * 16, cudaMemcpyHostToDevice, e->stream));
        d_win = e->windows.as<int32_t>();
    }
    B200X_TRY(sweep(e, B200X_MASK_OCCLUDE, n, d_win, occlusion_value, nullptr, false, e->prob.as<float>(), max_range, base_prob != nullptr));
    const cudaMemcpyKind back = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    B200X_CUDA_TRY(cudaMemcpyAsync(prob, e->prob.p, n * sizeof(float), back, e->stream));
    if (base_prob) B200X_CUDA_TRY(cudaMemcpyAsync(base_prob, e->prob.as<float>() + n, sizeof(float), back, e->stream));
    B200X_CUDA_TRY(cudaStreamSynchronize(e->stream));
    return B200X_OK;
}
}  // namespace

extern "C" int b200x_engine_fbp_sweep(b200x_engine* e, const float* gains, int n, int normalize_loudness, int on_device,
                                      float* prob) {
    B200X_TRY(check_ready(e, true));
    if (n == 0) return B200X_OK;
    B200X_REQUIRE(gains && prob && n > 0, "fbp_sweep: bad argument");
    B200X_TRY(ensure_prob(e, n));
    const float* d_g = gains;
    if (!on_device) {
        const size_t bytes = static_cast<size_t>(n) * b200x_engine::n_freq * sizeof(float);
        B200X_TRY(ensure_grow(e->gains, bytes));
        B200X_CUDA_TRY(cudaMemcpyAsync(e->gains.p, gains, bytes, cudaMemcpyHostToDevice, e->stream));
        d_g = e->gains.as<float>();
    }
    if (normalize_loudness) B200X_TRY(ensure_ref_rms(e));
    B200X_TRY(sweep(e, B200X_MASK_BAND_GAIN, n, nullptr, 0.f, d_g, normalize_loudness != 0, e->prob.as<float>()));
    B200X_CUDA_TRY(cudaMemcpyAsync(prob, e->prob.p, n * sizeof(float), on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, e->stream));
    B200X_CUDA_TRY(cudaStreamSynchronize(e->stream));
    return B200X_OK;
}

// FBP over a batch of equal-length tracks (BASELINE configs[2]: 64 tracks x the high_resolution bank): the band copies of
// as many tracks as fit one chunk go through ONE iSTFT launch and ONE classifier forward, the tracks' baselines through a
// second forward - instead of 13 + 1 copies per launch.  Per-copy arithmetic is that of fbp_sweep / predict, bit for bit.
extern "C" int b200x_engine_fbp_sweep_tracks(b200x_engine* e, const float* waves, int n_tracks, int64_t n_samples,
                                             const float* gains, int n_bands, int normalize_loudness, float* base_prob,
                                             float* prob) {
    B200X_TRY(check_ready(e, false));
    if (n_tracks == 0) return B200X_OK;
    B200X_REQUIRE(waves && gains && base_prob && prob && n_tracks > 0 && n_bands > 0, "fbp_sweep_tracks: bad argument");
    B200X_REQUIRE(n_samples >= 2048 && n_samples <= e->max_samples, "fbp_sweep_tracks: n_samples=%lld outside [2048, %lld]",
                  (long long)n_samples, (long long)e->max_samples);
    B200X_REQUIRE(n_bands <= e->C, "fbp_sweep_tracks: %d bands do not fit a chunk of %d copies", n_bands, e->C);
    const b200x_model_config& c = e->cfg;
    const int n_time = 1 + static_cast<int>(n_samples / c.hop_length);
    const int64_t out_len = static_cast<int64_t>(c.hop_length) * (n_time - 1);
    // when the iSTFT output is as long as the track (n_samples a multiple of the hop) the baselines ride in the SAME forward as
    // the band copies (one more copy per track, left unscaled); otherwise they take a forward of their own
    const bool joint = (out_len == n_samples) && (n_bands + 1 <= e->C);
    const int G = std::max(1, std::min(n_tracks, e->C / (n_bands + (joint ? 1 : 0))));   // tracks per group
    const int64_t track_stride = static_cast<int64_t>(n_time) * b200x_engine::s_stride;   // complex values per spectrogram
    B200X_TRY(ensure_grow(e->S_multi, static_cast<size_t>(G) * track_stride * 2 * sizeof(float)));
    B200X_TRY(ensure_grow(e->stems, static_cast<size_t>(G) * n_samples * sizeof(float)));
    B200X_TRY(ensure_grow(e->gains, static_cast<size_t>(G) * n_bands * b200x_engine::n_freq * sizeof(float)));
    B200X_TRY(ensure_grow(e->ref_arr, static_cast<size_t>(e->C) * sizeof(double)));
    B200X_TRY(ensure_prob(e, G * (n_bands + 1)));
    // the band table repeats for every track of a group
    for (int g = 0; g < G; ++g)
        B200X_CUDA_TRY(cudaMemcpyAsync(e->gains.as<float>() + static_cast<size_t>(g) * n_bands * b200x_engine::n_freq, gains,
                                       static_cast<size_t>(n_bands) * b200x_engine::n_freq * sizeof(float), cudaMemcpyHostToDevice, e->stream));
    e->baseline_valid = false;
    for (int t0 = 0; t0 < n_tracks; t0 += G) {
        const int g_n = std::min(G, n_tracks - t0);
        const int m = g_n * n_bands;
        float* d_waves = e->stems.as<float>();
        B200X_CUDA_TRY(cudaMemcpyAsync(d_waves, waves + static_cast<size_t>(t0) * n_samples, static_cast<size_t>(g_n) * n_samples * sizeof(float),
                                       cudaMemcpyHostToDevice, e->stream));
        for (int g = 0; g < g_n; ++g) {
            B200X_TRY(b200x_stft(d_waves + static_cast<size_t>(g) * n_samples, n_samples, c.n_fft, c.hop_length, 0,
                                 e->S_multi.as<float>() + static_cast<size_t>(g) * track_stride * 2, b200x_engine::s_stride, e->stream));
            e->launches += 1;
        }
        double* sumsq = nullptr;
        if (normalize_loudness) {
            sumsq = e->sumsq.as<double>();
            B200X_CUDA_TRY(cudaMemsetAsync(sumsq, 0, (m + g_n) * sizeof(double), e->stream));
            B200X_TRY(b200x_wave_rms(d_waves, n_samples, n_samples, g_n, n_bands, e->ref_arr.as<double>(), e->stream));
            e->launches += 1;
            if (joint) {                                  // reference level -1: the baseline copies are not rescaled
                std::vector<double> neg(g_n, -1.0);
                B200X_CUDA_TRY(cudaMemcpyAsync(e->ref_arr.as<double>() + m, neg.data(), g_n * sizeof(double), cudaMemcpyHostToDevice, e->stream));
                B200X_CUDA_TRY(cudaStreamSynchronize(e->stream));
            }
        }
        TIMED(KC_ISTFT, b200x_istft_masked_tracks(e->S_multi.p, b200x_engine::s_stride, n_time, m, n_bands, track_stride, B200X_MASK_BAND_GAIN,
                                            nullptr, 0.f, e->gains.as<float>(), e->y.as<float>(), e->y_stride, sumsq, nullptr, 0, e->stream));
        e->launches += 1;
        if (joint)                                        // baselines = the tracks themselves (dsp_band_ops.py:544), rows m .. m + g_n
            B200X_CUDA_TRY(cudaMemcpy2DAsync(e->y.as<float>() + static_cast<size_t>(m) * e->y_stride, e->y_stride * sizeof(float), d_waves,
                                             n_samples * sizeof(float), n_samples * sizeof(float), g_n, cudaMemcpyDeviceToDevice, e->stream));
        e->ref_arr_cur = normalize_loudness ? e->ref_arr.as<double>() : nullptr;
        const int rc = forward_chunk(e, m + (joint ? g_n : 0), out_len, sumsq, out_len, e->prob.as<float>(), e->logit.as<float>());
        e->ref_arr_cur = nullptr;
        B200X_TRY(rc);
        B200X_CUDA_TRY(cudaMemcpyAsync(prob + static_cast<size_t>(t0) * n_bands, e->prob.p, static_cast<size_t>(m) * sizeof(float),
                                       cudaMemcpyDeviceToHost, e->stream));
        if (joint) {
            B200X_CUDA_TRY(cudaMemcpyAsync(base_prob + t0, e->prob.as<float>() + m, static_cast<size_t>(g_n) * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
        } else {
            // one forward for all the baselines of the group
            B200X_CUDA_TRY(cudaMemsetAsync(e->y.p, 0, static_cast<size_t>(g_n) * e->y_stride * sizeof(float), e->stream));
            B200X_CUDA_TRY(cudaMemcpy2DAsync(e->y.p, e->y_stride * sizeof(float), d_waves, n_samples * sizeof(float), n_samples * sizeof(float),
                                             g_n, cudaMemcpyDeviceToDevice, e->stream));
            B200X_TRY(forward_chunk(e, g_n, n_samples, nullptr, 0, e->prob.as<float>(), e->logit.as<float>()));
            B200X_CUDA_TRY(cudaMemcpyAsync(base_prob + t0, e->prob.p, static_cast<size_t>(g_n) * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
        }
        B200X_CUDA_TRY(cudaStreamSynchronize(e->stream));
        if (t0 + g_n == n_tracks) {
            // the LAST track of the batch becomes the engine's current track (shape queries, band_map, spectrogram)
            const int g = g_n - 1;
            B200X_CUDA_TRY(cudaMemcpyAsync(e->wave.p, d_waves + static_cast<size_t>(g) * n_samples, n_samples * sizeof(float), cudaMemcpyDeviceToDevice, e->stream));
            B200X_CUDA_TRY(cudaMemcpyAsync(e->S.p, e->S_multi.as<float>() + static_cast<size_t>(g) * track_stride * 2,
                                           static_cast<size_t>(track_stride) * 2 * sizeof(float), cudaMemcpyDeviceToDevice, e->stream));
            e->L = n_samples;
            e->n_time = n_time;
            e->mel_track_frames = 0;
            e->ref_rms = -1.0;
            // the tail of every y row beyond hop * (n_time - 1) must read as zero padding again for the occlusion path
            B200X_CUDA_TRY(cudaMemsetAsync(e->y.p, 0, e->y.bytes, e->stream));
            B200X_CUDA_TRY(cudaStreamSynchronize(e->stream));
        }
    }
    return B200X_OK;
}

extern "C" int b200x_engine_stem_sweep(b200x_engine* e, const float* stems, int n_stems, int64_t n_samples,
                                       const uint8_t* masks, int n, int on_device, float* prob) {
    B200X_TRY(check_ready(e, false));
    if (n == 0) return B200X_OK;
    B200X_REQUIRE(stems && masks && prob && n > 0 && n_stems > 0, "stem_sweep: bad argument");
    B200X_REQUIRE(n_samples > 1024 && n_samples <= e->max_samples, "stem_sweep: n_samples out of range");
    B200X_TRY(ensure_prob(e, n));
    const float* d_st = stems;
    const uint8_t* d_mk = masks;
    if (!on_device) {
        const size_t sb = static_cast<size_t>(n_stems) * n_samples * sizeof(float);
        B200X_TRY(ensure_grow(e->stems, sb));
        B200X_TRY(ensure_grow(e->masks, static_cast<size_t>(n) * n_stems));
        B200X_CUDA_TRY(cudaMemcpyAsync(e->stems.p, stems, sb, cudaMemcpyHostToDevice, e->stream));
        B200X_CUDA_TRY(cudaMemcpyAsync(e->masks.p, masks, static_cast<size_t>(n) * n_stems, cudaMemcpyHostToDevice, e->stream));
        d_st = e->stems.as<float>();
        d_mk = e->masks.as<uint8_t>();
    }
    for (int c0 = 0; c0 < n; c0 += e->C) {
        const int m = std::min(e->C, n - c0);
        B200X_TRY(b200x_mix_stems(d_st, n_samples, n_stems, d_mk + static_cast<size_t>(c0) * n_stems, m, e->y.as<float>(), e->y_stride, e->stream));
        e->launches += 1;
        B200X_TRY(forward_chunk(e, m, n_samples, nullptr, 0, e->prob.as<float>() + c0, e->logit.as<float>() + c0));
    }
    B200X_CUDA_TRY(cudaMemcpyAsync(prob, e->prob.p, n * sizeof(float), on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, e->stream));
    B200X_CUDA_TRY(cudaStreamSynchronize(e->stream));
    return B200X_OK;
}

namespace {
// perturbed audio to the host.  seg_stride > 0: only the window's own time span [t0*hop, t0*hop + max(1,(t1-t0)*hop))
// clipped to the iSTFT length is copied (row i at audio_host + i*seg_stride, valid length in seg_len[i]).
int audio_out(b200x_engine* e, int mode, const int32_t* windows, const float* gains, int n, float occ_value, float* audio_host,
              int64_t seg_stride, int64_t* seg_len) {
    const int64_t out_len = static_cast<int64_t>(e->cfg.hop_length) * (e->n_time - 1);
    if (windows) {
        B200X_TRY(ensure_grow(e->windows, static_cast<size_t>(n) * 16));
        B200X_CUDA_TRY(cudaMemcpyAsync(e->windows.p, windows, static_cast<size_t>(n) * 16, cudaMemcpyHostToDevice, e->stream));
    }
    if (gains) {
        const size_t bytes = static_cast<size_t>(n) * b200x_engine::n_freq * sizeof(float);
        B200X_TRY(ensure_grow(e->gains, bytes));
        B200X_CUDA_TRY(cudaMemcpyAsync(e->gains.p, gains, bytes, cudaMemcpyHostToDevice, e->stream));
    }
    // window segments: only the hops of the window's own time span are synthesised (frame range [t0, t1] -> hops t0 .. t1 + 2)
    const bool seg_sparse = seg_stride > 0 && windows != nullptr;
    int max_range = 0;
    if (seg_sparse) {
        std::vector<int32_t> rg(static_cast<size_t>(n) * 2);
        for (int i = 0; i < n; ++i) {
            rg[2 * i] = windows[4 * i];
            rg[2 * i + 1] = std::max(windows[4 * i + 1], windows[4 * i] + 1);
            max_range = std::max(max_range, rg[2 * i + 1] - rg[2 * i]);
        }
        B200X_TRY(ensure_grow(e->ranges, rg.size() * sizeof(int32_t)));
        B200X_CUDA_TRY(cudaMemcpyAsync(e->ranges.p, rg.data(), rg.size() * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
        B200X_CUDA_TRY(cudaStreamSynchronize(e->stream));      // rg is a stack-lifetime staging buffer
    }
    for (int c0 = 0; c0 < n; c0 += e->C) {
        const int m = std::min(e->C, n - c0);
        TIMED(KC_ISTFT, b200x_istft_masked(e->S.p, b200x_engine::s_stride, e->n_time, m, mode, windows ? e->windows.as<int32_t>() + 4 * c0 : nullptr,
                                     occ_value, gains ? e->gains.as<float>() + static_cast<size_t>(c0) * b200x_engine::n_freq : nullptr,
                                     e->y.as<float>(), e->y_stride, nullptr, seg_sparse ? e->ranges.as<int32_t>() + 2 * c0 : nullptr,
                                     max_range, e->stream));
        e->launches += 1;
        if (seg_stride > 0) {
            for (int i = 0; i < m; ++i) {
                const int32_t* w = windows + 4 * (c0 + i);
                const int64_t want = std::max<int64_t>(1, static_cast<int64_t>(w[1] - w[0]) * e->cfg.hop_length);
                const int64_t start = std::min<int64_t>(static_cast<int64_t>(w[0]) * e->cfg.hop_length, out_len);
                const int64_t len = std::min<int64_t>(std::min(start + want, out_len) - start, seg_stride);
                if (seg_len) seg_len[c0 + i] = len;
                if (len > 0)
                    B200X_CUDA_TRY(cudaMemcpyAsync(audio_host + static_cast<size_t>(c0 + i) * seg_stride,
                                                   e->y.as<float>() + static_cast<size_t>(i) * e->y_stride + start, len * sizeof(float),
                                                   cudaMemcpyDeviceToHost, e->stream));
            }
        } else {
            B200X_CUDA_TRY(cudaMemcpy2DAsync(audio_host + static_cast<size_t>(c0) * out_len, out_len * sizeof(float), e->y.p,
                                             e->y_stride * sizeof(float), out_len * sizeof(float), m, cudaMemcpyDeviceToHost, e->stream));
        }
        // the next chunk overwrites y: the copies above must have drained first
        B200X_CUDA_TRY(cudaStreamSynchronize(e->stream));
    }
    return B200X_OK;
}
}  // namespace

extern "C" int b200x_engine_window_audio(b200x_engine* e, const int32_t* windows, int n, float* audio_host,
                                         int64_t seg_stride, int64_t* seg_len) {
    B200X_TRY(check_ready(e, true));
    if (n == 0) return B200X_OK;
    B200X_REQUIRE(windows && audio_host && n > 0 && seg_stride > 0, "window_audio: bad argument");
    return audio_out(e, B200X_MASK_KEEP_ONLY, windows, nullptr, n, 0.f, audio_host, seg_stride, seg_len);
}

extern "C" int b200x_engine_occluded_audio(b200x_engine* e, const int32_t* windows, int n, float occlusion_value,
                                           float* audio_host) {
    B200X_TRY(check_ready(e, true));
    if (n == 0) return B200X_OK;
    B200X_REQUIRE(windows && audio_host && n > 0, "occluded_audio: bad argument");
    return audio_out(e, B200X_MASK_OCCLUDE, windows, nullptr, n, occlusion_value, audio_host, 0, nullptr);
}

extern "C" int b200x_engine_band_audio(b200x_engine* e, const float* gains, int n, float* audio_host) {
    B200X_TRY(check_ready(e, true));
    if (n == 0) return B200X_OK;
    B200X_REQUIRE(gains && audio_host && n > 0, "band_audio: bad argument");
    return audio_out(e, B200X_MASK_BAND_GAIN, nullptr, gains, n, 0.f, audio_host, 0, nullptr);
}

namespace {
// RISE rows for the iSTFT load stage: (seed, mask index, keep threshold, 0) per perturbed copy
std::vector<int32_t> rise_rows(int first_mask, int n, uint32_t seed, double keep_probability) {
    std::vector<int32_t> rows(static_cast<size_t>(n) * 4);
    const uint32_t thr = rise_threshold(keep_probability);
    for (int i = 0; i < n; ++i) {
        rows[4 * i] = static_cast<int32_t>(seed);
        rows[4 * i + 1] = first_mask + i;
        rows[4 * i + 2] = static_cast<int32_t>(thr);
        rows[4 * i + 3] = 0;
    }
    return rows;
}
}  // namespace

extern "C" int b200x_engine_rise_sweep(b200x_engine* e, int first_mask, int n, uint32_t seed, double keep_probability,
                                       int on_device, float* prob) {
    B200X_TRY(check_ready(e, true));
    if (n == 0) return B200X_OK;
    B200X_REQUIRE(prob && n > 0 && first_mask >= 0, "rise_sweep: bad argument");
    B200X_REQUIRE(keep_probability > 0.0 && keep_probability <= 1.0, "rise_sweep: keep probability %g outside (0, 1]", keep_probability);
    B200X_TRY(ensure_prob(e, n));
    const std::vector<int32_t> rows = rise_rows(first_mask, n, seed, keep_probability);
    B200X_TRY(ensure_grow(e->windows, rows.size() * sizeof(int32_t)));
    B200X_CUDA_TRY(cudaMemcpyAsync(e->windows.p, rows.data(), rows.size() * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
    B200X_CUDA_TRY(cudaStreamSynchronize(e->stream));          // rows is a stack-lifetime staging buffer
    B200X_TRY(sweep(e, B200X_MASK_RANDOM_KEEP, n, e->windows.as<int32_t>(), 0.f, nullptr, false, e->prob.as<float>()));
    B200X_CUDA_TRY(cudaMemcpyAsync(prob, e->prob.p, n * sizeof(float), on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, e->stream));
    B200X_CUDA_TRY(cudaStreamSynchronize(e->stream));
    return B200X_OK;
}

extern "C" int b200x_engine_rise_audio(b200x_engine* e, int first_mask, int n, uint32_t seed, double keep_probability,
                                       float* audio_host) {
    B200X_TRY(check_ready(e, true));
    if (n == 0) return B200X_OK;
    B200X_REQUIRE(audio_host && n > 0 && first_mask >= 0, "rise_audio: bad argument");
    const std::vector<int32_t> rows = rise_rows(first_mask, n, seed, keep_probability);
    return audio_out(e, B200X_MASK_RANDOM_KEEP, rows.data(), nullptr, n, 0.f, audio_host, 0, nullptr);
}

extern "C" int b200x_engine_rise_map(b200x_engine* e, const double* pred, int n, uint32_t seed, double keep_probability,
                                     double* map_host) {
    B200X_TRY(check_ready(e, true));
    B200X_REQUIRE(map_host && n >= 0 && (n == 0 || pred), "rise_map: bad argument");
    const size_t cells = static_cast<size_t>(b200x_engine::n_freq) * e->n_time;
    B200X_TRY(ensure_grow(e->map, cells * sizeof(double)));
    B200X_TRY(ensure_grow(e->delta, static_cast<size_t>(std::max(n, 1)) * sizeof(double)));
    if (n > 0) B200X_CUDA_TRY(cudaMemcpyAsync(e->delta.p, pred, static_cast<size_t>(n) * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    B200X_TRY(b200x_rise_map(e->delta.as<double>(), n, seed, keep_probability, b200x_engine::n_freq, e->n_time, e->map.as<double>(), e->stream));
    e->launches += 1;
    B200X_CUDA_TRY(cudaMemcpyAsync(map_host, e->map.p, cells * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    B200X_CUDA_TRY(cudaStreamSynchronize(e->stream));
    return B200X_OK;
}

extern "C" int b200x_engine_saliency_map(b200x_engine* e, const int32_t* windows, const double* delta, int n,
                                         double* map_host) {
    B200X_TRY(check_ready(e, true));
    B200X_REQUIRE(map_host && n >= 0 && (n == 0 || (windows && delta)), "saliency_map: bad argument");
    const size_t cells = static_cast<size_t>(b200x_engine::n_freq) * e->n_time;
    B200X_TRY(ensure_grow(e->map, cells * sizeof(double)));
    B200X_TRY(ensure_grow(e->windows, static_cast<size_t>(std::max(n, 1)) * 16));
    B200X_TRY(ensure_grow(e->delta, static_cast<size_t>(std::max(n, 1)) * sizeof(double)));
    if (n > 0) {
        B200X_CUDA_TRY(cudaMemcpyAsync(e->windows.p, windows, static_cast<size_t>(n) * 16, cudaMemcpyHostToDevice, e->stream));
        B200X_CUDA_TRY(cudaMemcpyAsync(e->delta.p, delta, static_cast<size_t>(n) * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    }
    B200X_TRY(b200x_saliency_reduce(e->windows.as<int32_t>(), e->delta.as<double>(), n, b200x_engine::n_freq, e->n_time, e->map.as<double>(), e->stream));
    e->launches += 1;
    B200X_CUDA_TRY(cudaMemcpyAsync(map_host, e->map.p, cells * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    B200X_CUDA_TRY(cudaStreamSynchronize(e->stream));
    return B200X_OK;
}

extern "C" int b200x_engine_band_map(b200x_engine* e, const int32_t* band_rows, const double* delta, int n,
                                     double* map_host) {
    B200X_TRY(check_ready(e, true));
    B200X_REQUIRE(map_host && n >= 0 && (n == 0 || (band_rows && delta)), "band_map: bad argument");
    const size_t cells = static_cast<size_t>(b200x_engine::n_freq) * e->n_time;
    B200X_TRY(ensure_grow(e->map, cells * sizeof(double)));
    B200X_TRY(ensure_grow(e->windows, static_cast<size_t>(std::max(n, 1)) * 8));
    B200X_TRY(ensure_grow(e->delta, static_cast<size_t>(std::max(n, 1)) * sizeof(double)));
    if (n > 0) {
        B200X_CUDA_TRY(cudaMemcpyAsync(e->windows.p, band_rows, static_cast<size_t>(n) * 8, cudaMemcpyHostToDevice, e->stream));
        B200X_CUDA_TRY(cudaMemcpyAsync(e->delta.p, delta, static_cast<size_t>(n) * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    }
    B200X_TRY(b200x_band_map(e->windows.as<int32_t>(), e->delta.as<double>(), n, b200x_engine::n_freq, e->n_time, e->map.as<double>(), e->stream));
    e->launches += 1;
    B200X_CUDA_TRY(cudaMemcpyAsync(map_host, e->map.p, cells * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    B200X_CUDA_TRY(cudaStreamSynchronize(e->stream));
    return B200X_OK;
}

extern "C" int b200x_engine_rank(b200x_engine* e, const double* values, int n, int mode, int32_t* order_host) {
    B200X_REQUIRE(e != nullptr, "rank: engine is NULL");
    B200X_CUDA_TRY(cudaSetDevice(e->device));
    if (n == 0) return B200X_OK;
    B200X_REQUIRE(values && order_host && n > 0, "rank: bad argument");
    B200X_TRY(ensure_grow(e->delta, static_cast<size_t>(n) * sizeof(double)));
    B200X_TRY(ensure_grow(e->order, static_cast<size_t>(n) * sizeof(int32_t)));
    B200X_CUDA_TRY(cudaMemcpyAsync(e->delta.p, values, static_cast<size_t>(n) * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    B200X_TRY(b200x_rank(e->delta.as<double>(), n, mode, e->order.as<int32_t>(), e->stream));
    e->launches += 1;
    B200X_CUDA_TRY(cudaMemcpyAsync(order_host, e->order.p, static_cast<size_t>(n) * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
    B200X_CUDA_TRY(cudaStreamSynchronize(e->stream));
    return B200X_OK;
}

// Track loader front: resample a decoded track on the device (host buffers in / out; works without weights)
extern "C" int b200x_engine_resample(b200x_engine* e, const float* x_host, int64_t n_in, int up, int down, const double* h_host,
                                     int h_len, float* y_host, int64_t n_out) {
    B200X_REQUIRE(e && x_host && h_host && y_host && n_in > 0 && n_out > 0, "resample: bad argument");
    B200X_CUDA_TRY(cudaSetDevice(e->device));
    DevBuf dx, dh, dy;
    int st = dx.alloc(static_cast<size_t>(n_in) * sizeof(float));
    if (st == B200X_OK) st = dh.alloc(static_cast<size_t>(h_len) * sizeof(double));
    if (st == B200X_OK) st = dy.alloc(static_cast<size_t>(n_out) * sizeof(float));
    auto run = [&]() -> int {
        B200X_CUDA_TRY(cudaMemcpyAsync(dx.p, x_host, static_cast<size_t>(n_in) * sizeof(float), cudaMemcpyHostToDevice, e->stream));
        B200X_CUDA_TRY(cudaMemcpyAsync(dh.p, h_host, static_cast<size_t>(h_len) * sizeof(double), cudaMemcpyHostToDevice, e->stream));
        B200X_TRY(b200x_resample_poly(dx.as<float>(), n_in, up, down, dh.as<double>(), h_len, dy.as<float>(), n_out, e->stream));
        e->launches += 1;
        B200X_CUDA_TRY(cudaMemcpyAsync(y_host, dy.p, static_cast<size_t>(n_out) * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
        B200X_CUDA_TRY(cudaStreamSynchronize(e->stream));
        return B200X_OK;
    };
    if (st == B200X_OK) st = run();
    dx.release(); dh.release(); dy.release();
    return st;
}

// importance map over an arbitrary [n_freq][n_time] grid (the mel variant's map has n_mels rows)
extern "C" int b200x_engine_saliency_map_shape(b200x_engine* e, const int32_t* windows, const double* delta, int n, int n_freq,
                                               int n_time, double* map_host) {
    B200X_TRY(check_ready(e, false));
    B200X_REQUIRE(map_host && n >= 0 && n_freq > 0 && n_time > 0 && (n == 0 || (windows && delta)), "saliency_map_shape: bad argument");
    const size_t cells = static_cast<size_t>(n_freq) * n_time;
    B200X_TRY(ensure_grow(e->map, cells * sizeof(double)));
    B200X_TRY(ensure_grow(e->windows, static_cast<size_t>(std::max(n, 1)) * 16));
    B200X_TRY(ensure_grow(e->delta, static_cast<size_t>(std::max(n, 1)) * sizeof(double)));
    if (n > 0) {
        B200X_CUDA_TRY(cudaMemcpyAsync(e->windows.p, windows, static_cast<size_t>(n) * 16, cudaMemcpyHostToDevice, e->stream));
        B200X_CUDA_TRY(cudaMemcpyAsync(e->delta.p, delta, static_cast<size_t>(n) * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    }
    B200X_TRY(b200x_saliency_reduce(e->windows.as<int32_t>(), e->delta.as<double>(), n, n_freq, n_time, e->map.as<double>(), e->stream));
    e->launches += 1;
    B200X_CUDA_TRY(cudaMemcpyAsync(map_host, e->map.p, cells * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    B200X_CUDA_TRY(cudaStreamSynchronize(e->stream));
    return B200X_OK;
}

// ---------------------------------------------------------------------------------------------------- mel-domain variant
extern "C" int b200x_engine_set_mel_basis(b200x_engine* e, int n_mels, const float* basis, const float* pinv, float step) {
    B200X_TRY(check_ready(e, false));
    B200X_REQUIRE(basis && pinv && n_mels > 0 && n_mels <= 1024 && step > 0.f, "set_mel_basis: bad argument");
    const int F = b200x_engine::n_freq;
    // sparse structure of the filterbank: the bins of every filter, and the (at most two, adjacent) filters of every bin
    std::vector<int32_t> bin_range(2 * static_cast<size_t>(n_mels)), bin_first(F, 0);
    std::vector<float> bin_w(2 * static_cast<size_t>(F), 0.f), pinv_t(static_cast<size_t>(n_mels) * F);
    for (int i = 0; i < n_mels; ++i) {
        int lo = F, hi = 0;
        for (int k = 0; k < F; ++k)
            if (basis[static_cast<size_t>(i) * F + k] != 0.f) { lo = std::min(lo, k); hi = std::max(hi, k + 1); }
        if (lo >= hi) { lo = 0; hi = 0; }                // empty filter (narrower than the bin spacing)
        bin_range[2 * i] = lo; bin_range[2 * i + 1] = hi;
    }
    for (int k = 0; k < F; ++k) {
        int first = -1, count = 0, last = -1;
        for (int i = 0; i < n_mels; ++i)
            if (basis[static_cast<size_t>(i) * F + k] != 0.f) { if (first < 0) first = i; last = i; ++count; }
        B200X_REQUIRE(count <= 2 && (count < 2 || last == first + 1),
                      "set_mel_basis: bin %d feeds %d filters (%d..%d); triangular filterbanks feed at most two adjacent ones", k, count, first, last);
        if (first < 0) first = 0;
        bin_first[k] = first;
        bin_w[2 * k] = basis[static_cast<size_t>(first) * F + k];
        bin_w[2 * k + 1] = first + 1 < n_mels ? basis[static_cast<size_t>(first + 1) * F + k] : 0.f;
    }
    for (int k = 0; k < F; ++k)
        for (int i = 0; i < n_mels; ++i) pinv_t[static_cast<size_t>(i) * F + k] = pinv[static_cast<size_t>(k) * n_mels + i];
    B200X_CUDA_TRY(cudaStreamSynchronize(e->stream));
    B200X_TRY(upload(e->mel_basis, basis, static_cast<size_t>(n_mels) * F * sizeof(float)));
    B200X_TRY(upload(e->mel_bin_range, bin_range.data(), bin_range.size() * sizeof(int32_t)));
    B200X_TRY(upload(e->mel_bin_first, bin_first.data(), bin_first.size() * sizeof(int32_t)));
    B200X_TRY(upload(e->mel_bin_w, bin_w.data(), bin_w.size() * sizeof(float)));
    B200X_TRY(upload(e->mel_pinv_t, pinv_t.data(), pinv_t.size() * sizeof(float)));
    e->mel_n = n_mels;
    e->mel_step = step;
    e->mel_track_frames = 0;
    return B200X_OK;
}

namespace {
// power mel spectrogram of the current track (once per track and basis)
int ensure_mel_track(b200x_engine* e) {
    B200X_REQUIRE(e->mel_n > 0, "mel: no filterbank uploaded (b200x_engine_set_mel_basis)");
    if (e->mel_track_frames == e->n_time) return B200X_OK;
    B200X_TRY(ensure_grow(e->mel_track, static_cast<size_t>(e->n_time) * e->mel_n * sizeof(float)));
    B200X_TRY(b200x_mel_power(e->S.p, b200x_engine::s_stride, e->n_time, e->mel_n, e->mel_basis.as<float>(), e->mel_bin_range.as<int32_t>(),
                              e->mel_track.as<float>(), e->stream));
    e->launches += 1;
    e->mel_track_frames = e->n_time;
    e->mel_mag_iter = -1;
    return B200X_OK;
}

int nnls_call(b200x_engine* e, int copies, int mode, const int32_t* d_windows, float occ, const float* d_gains, int nnls_iter,
              const int32_t* d_frames, int max_range, float* d_mag, int64_t mag_copy_stride) {
    B200X_TRY(b200x_mel_nnls(e->mel_track.as<float>(), e->n_time, e->mel_n, copies, mode, d_windows, occ, d_gains, e->mel_basis.as<float>(),
                             e->mel_bin_range.as<int32_t>(), e->mel_bin_first.as<int32_t>(), e->mel_bin_w.as<float>(), e->mel_pinv_t.as<float>(),
                             e->mel_step, nnls_iter, d_frames, max_range, d_mag, mag_copy_stride, e->stream));
    e->launches += 1;
    return B200X_OK;
}
}  // namespace

extern "C" int b200x_engine_mel_spectrogram(b200x_engine* e, float* mel_host) {
    B200X_TRY(check_ready(e, true));
    B200X_REQUIRE(mel_host != nullptr, "mel_spectrogram: NULL output");
    B200X_TRY(ensure_mel_track(e));
    std::vector<float> tmp(static_cast<size_t>(e->n_time) * e->mel_n);
    B200X_CUDA_TRY(cudaMemcpyAsync(tmp.data(), e->mel_track.p, tmp.size() * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
    B200X_CUDA_TRY(cudaStreamSynchronize(e->stream));
    for (int t = 0; t < e->n_time; ++t)                       // [frame][mel] on the device -> librosa's [mel][frame]
        for (int i = 0; i < e->mel_n; ++i) mel_host[static_cast<size_t>(i) * e->n_time + t] = tmp[static_cast<size_t>(t) * e->mel_n + i];
    return B200X_OK;
}

// The mel-variant hot loop (src/spectrogram_explainability.py:663-703 with spec_type == 'mel', and the builder's FBP-mel):
// for every perturbed copy  y = griffinlim(sqrt(nnls(A, perturbed mel))), trimmed / padded to len(track), prob = predict(y).
// mode: B200X_MASK_OCCLUDE / KEEP_ONLY (windows int32 [n][4] over (frame, mel bin)) or BAND_GAIN (gains float [n][n_mels]).
// Copy i starts Griffin-Lim from the phases of index first_index + i.  audio_host (nullable): float [n][hop * (n_time - 1)].
extern "C" int b200x_engine_mel_sweep(b200x_engine* e, int mode, const int32_t* windows, const float* gains, int n, float occlusion_value,
                                      int n_iter, int nnls_iter, uint32_t seed, int first_index, float momentum, float* prob,
                                      float* audio_host) {
    B200X_TRY(check_ready(e, true));
    if (n == 0) return B200X_OK;
    B200X_REQUIRE(n > 0 && n_iter >= 0 && nnls_iter >= 0 && first_index >= 0 && (prob || audio_host), "mel_sweep: bad argument");
    B200X_REQUIRE(mode == B200X_MASK_OCCLUDE || mode == B200X_MASK_KEEP_ONLY || mode == B200X_MASK_BAND_GAIN || mode == B200X_MASK_NONE,
                  "mel_sweep: bad mode %d", mode);
    B200X_REQUIRE((mode != B200X_MASK_OCCLUDE && mode != B200X_MASK_KEEP_ONLY) || windows, "mel_sweep: windows missing");
    B200X_REQUIRE(mode != B200X_MASK_BAND_GAIN || gains, "mel_sweep: gains missing");
    B200X_TRY(ensure_mel_track(e));
    const int T = e->n_time, NM = e->mel_n;
    const int64_t spec_copy = static_cast<int64_t>(T) * b200x_engine::s_stride;      // elements per copy (float or float2)
    const int64_t out_len = static_cast<int64_t>(e->cfg.hop_length) * (T - 1);
    const bool window_mode = mode == B200X_MASK_OCCLUDE || mode == B200X_MASK_KEEP_ONLY;
    // Griffin-Lim workspace: 28 bytes per cell and copy (|STFT|, C, two rebuilt spectra): chunks of at most 16 copies
    const int G = std::min(std::min(e->C, 16), n);
    if (e->gl_copies < G) {
        B200X_TRY(e->gl_mag.alloc(static_cast<size_t>(G) * spec_copy * sizeof(float)));
        B200X_TRY(e->gl_c.alloc(static_cast<size_t>(G) * spec_copy * 2 * sizeof(float)));
        B200X_TRY(e->gl_r0.alloc(static_cast<size_t>(G) * spec_copy * 2 * sizeof(float)));
        B200X_TRY(e->gl_r1.alloc(static_cast<size_t>(G) * spec_copy * 2 * sizeof(float)));
        e->gl_copies = G;
    }
    B200X_TRY(ensure_prob(e, n));
    std::vector<int32_t> frames(static_cast<size_t>(n) * 2);
    int max_range = 1;
    for (int i = 0; i < n; ++i) {
        int fa = 0, fb = T;
        if (window_mode) {
            const int32_t* w = windows + 4 * i;
            B200X_REQUIRE(w[0] >= 0 && w[0] <= w[1] && w[1] <= T && w[2] >= 0 && w[2] <= w[3] && w[3] <= NM,
                          "mel_sweep: window %d = (%d,%d,%d,%d) outside the %dx%d mel spectrogram", i, w[0], w[1], w[2], w[3], NM, T);
            fa = w[0]; fb = w[1];
        }
        frames[2 * i] = fa; frames[2 * i + 1] = fb;
        max_range = std::max(max_range, fb - fa);
    }
    B200X_TRY(ensure_grow(e->mel_frames, std::max<size_t>(frames.size(), 2) * sizeof(int32_t)));
    B200X_CUDA_TRY(cudaMemcpyAsync(e->mel_frames.p, frames.data(), frames.size() * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
    if (window_mode) {
        B200X_TRY(ensure_grow(e->windows, static_cast<size_t>(n) * 16));
        B200X_CUDA_TRY(cudaMemcpyAsync(e->windows.p, windows, static_cast<size_t>(n) * 16, cudaMemcpyHostToDevice, e->stream));
    }
    if (mode == B200X_MASK_BAND_GAIN) {
        B200X_TRY(ensure_grow(e->gains, static_cast<size_t>(n) * NM * sizeof(float)));
        B200X_CUDA_TRY(cudaMemcpyAsync(e->gains.p, gains, static_cast<size_t>(n) * NM * sizeof(float), cudaMemcpyHostToDevice, e->stream));
    }
    B200X_CUDA_TRY(cudaStreamSynchronize(e->stream));          // `frames` is a stack-lifetime staging buffer
    // frames outside an occlusion window keep the unperturbed track's NNLS solution (the NNLS is separable per frame)
    if (mode == B200X_MASK_OCCLUDE && e->mel_mag_iter != nnls_iter) {
        B200X_TRY(ensure_grow(e->mel_mag_base, static_cast<size_t>(spec_copy) * sizeof(float)));
        const int32_t all[2] = {0, T};
        B200X_TRY(ensure_grow(e->ranges, 2 * sizeof(int32_t)));
        B200X_CUDA_TRY(cudaMemcpyAsync(e->ranges.p, all, sizeof(all), cudaMemcpyHostToDevice, e->stream));
        B200X_CUDA_TRY(cudaStreamSynchronize(e->stream));
        B200X_TRY(nnls_call(e, 1, B200X_MASK_NONE, nullptr, 0.f, nullptr, nnls_iter, e->ranges.as<int32_t>(), T, e->mel_mag_base.as<float>(), spec_copy));
        e->mel_mag_iter = nnls_iter;
    }
    const float coef = momentum / (1.0f + momentum);
    for (int c0 = 0; c0 < n; c0 += G) {
        const int m = std::min(G, n - c0);
        float* mag = e->gl_mag.as<float>();
        if (mode == B200X_MASK_OCCLUDE) {
            for (int i = 0; i < m; ++i)
                B200X_CUDA_TRY(cudaMemcpyAsync(mag + static_cast<size_t>(i) * spec_copy, e->mel_mag_base.p, static_cast<size_t>(spec_copy) * sizeof(float),
                                               cudaMemcpyDeviceToDevice, e->stream));
        } else if (mode == B200X_MASK_KEEP_ONLY) {
            B200X_CUDA_TRY(cudaMemsetAsync(mag, 0, static_cast<size_t>(m) * spec_copy * sizeof(float), e->stream));
        }
        B200X_TRY(nnls_call(e, m, mode, window_mode ? e->windows.as<int32_t>() + 4 * c0 : nullptr, occlusion_value,
                            mode == B200X_MASK_BAND_GAIN ? e->gains.as<float>() + static_cast<size_t>(c0) * NM : nullptr, nnls_iter,
                            e->mel_frames.as<int32_t>() + 2 * c0, max_range, mag, spec_copy));
        B200X_TRY(b200x_gl_init(mag, spec_copy, e->gl_c.p, spec_copy, m, T, seed, first_index + c0, e->stream));
        e->launches += 1;
        for (int it = 0; it < n_iter; ++it) {
            void* rebuilt = (it & 1) ? e->gl_r1.p : e->gl_r0.p;
            void* tprev = (it & 1) ? e->gl_r0.p : e->gl_r1.p;
            TIMED(KC_ISTFT, b200x_istft_masked_tracks(e->gl_c.p, b200x_engine::s_stride, T, m, 1, spec_copy, B200X_MASK_NONE, nullptr, 0.f, nullptr,
                                                e->y.as<float>(), e->y_stride, nullptr, nullptr, 0, e->stream));
            TIMED(KC_OTHER, b200x_stft_batch(e->y.as<float>(), out_len, e->y_stride, m, rebuilt, b200x_engine::s_stride, spec_copy, e->stream));
            TIMED(KC_OTHER, b200x_gl_update(rebuilt, tprev, mag, spec_copy, e->gl_c.p, spec_copy, m, T, it == 0 ? 0.f : coef, e->stream));
            e->launches += 3;
        }
        TIMED(KC_ISTFT, b200x_istft_masked_tracks(e->gl_c.p, b200x_engine::s_stride, T, m, 1, spec_copy, B200X_MASK_NONE, nullptr, 0.f, nullptr,
                                            e->y.as<float>(), e->y_stride, nullptr, nullptr, 0, e->stream));
        e->launches += 1;
        if (audio_host != nullptr)
            B200X_CUDA_TRY(cudaMemcpy2DAsync(audio_host + static_cast<size_t>(c0) * out_len, out_len * sizeof(float), e->y.p,
                                             e->y_stride * sizeof(float), out_len * sizeof(float), m, cudaMemcpyDeviceToHost, e->stream));
        if (prob != nullptr) {
            // the reference trims / zero-pads the inverted audio to len(y) before predicting (:676-680); rows keep a zero tail
            if (e->L > out_len)
                B200X_CUDA_TRY(cudaMemset2DAsync(e->y.as<float>() + out_len, e->y_stride * sizeof(float), 0, (e->L - out_len) * sizeof(float), m, e->stream));
            B200X_TRY(forward_chunk(e, m, e->L, nullptr, 0, e->prob.as<float>() + c0, e->logit.as<float>() + c0));
        }
        B200X_CUDA_TRY(cudaStreamSynchronize(e->stream));       // audio_host rows / y are reused by the next chunk
    }
    if (prob != nullptr) {
        B200X_CUDA_TRY(cudaMemcpyAsync(prob, e->prob.p, static_cast<size_t>(n) * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
        B200X_CUDA_TRY(cudaStreamSynchronize(e->stream));
    }
    return B200X_OK;
}

extern "C" int b200x_engine_debug_buffer(b200x_engine* e, const char* name, void** d_ptr, int64_t* bytes) {
    B200X_REQUIRE(e && name && d_ptr, "debug_buffer: bad argument");
    const std::string n(name);
    DevBuf* b = nullptr;
    if (n == "y") b = &e->y; else if (n == "db") b = &e->db; else if (n == "img_t") b = &e->img_t;
    else if (n == "img_f") b = &e->img_f; else if (n == "x") b = &e->x; else if (n == "h") b = &e->h;
    else if (n == "qkv") b = &e->qkv; else if (n == "att") b = &e->att; else if (n == "hid") b = &e->hid;
    else if (n == "prob") b = &e->prob; else if (n == "logit") b = &e->logit; else if (n == "S") b = &e->S;
    else if (n == "wave") b = &e->wave;
    else return set_error(B200X_ERR_INVALID, "debug_buffer: unknown buffer '%s'", name);
    if (b->p == nullptr) return set_error(B200X_ERR_STATE, "debug_buffer: '%s' is not materialised in this configuration", name);
    *d_ptr = b->p;
    if (bytes) *bytes = static_cast<int64_t>(b->bytes);
    return B200X_OK;
}

extern "C" int b200x_engine_set_trace(b200x_engine* e, float* d_trace) {
    B200X_REQUIRE(e != nullptr, "set_trace: engine is NULL");
    e->trace = d_trace;
    return B200X_OK;
}

extern "C" int64_t b200x_engine_launch_count(b200x_engine* e) { return e ? e->launches : -1; }
extern "C" void* b200x_engine_stream(b200x_engine* e) { return e ? static_cast<void*>(e->stream) : nullptr; }
extern "C" int b200x_engine_synchronize(b200x_engine* e) {
    B200X_REQUIRE(e != nullptr, "synchronize: engine is NULL");
    B200X_CUDA_TRY(cudaStreamSynchronize(e->stream));
    return B200X_OK;
}

extern "C" int b200x_engine_set_alternate(b200x_engine* e, int enable) {
    if (e == nullptr) return set_error(B200X_ERR_INVALID, "engine is NULL");
    if (e->alternate != (enable != 0)) {
        e->alternate = enable != 0;
        for (auto& kv : e->graphs) {                  // the direction is baked into captured launches: drop the graphs
            if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
        }
        e->graphs.clear();
    }
    return B200X_OK;
}

extern "C" int b200x_engine_set_fused_layernorm(b200x_engine* e, int enable) {
    B200X_REQUIRE(e != nullptr, "set_fused_layernorm: engine is NULL");
    if (e->fuse_ln != (enable != 0)) {
        e->fuse_ln = enable != 0;
        for (auto& kv : e->graphs)                        // the launch sequence is baked into the captured graphs
            if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
        e->graphs.clear();
    }
    return B200X_OK;
}

extern "C" int b200x_engine_set_graphs(b200x_engine* e, int enable) {
    B200X_REQUIRE(e != nullptr, "set_graphs: engine is NULL");
    e->use_graphs = enable != 0;
    return B200X_OK;
}

extern "C" int b200x_engine_set_timing(b200x_engine* e, int enable) {
    B200X_REQUIRE(e != nullptr, "set_timing: engine is NULL");
    B200X_CUDA_TRY(cudaStreamSynchronize(e->stream));
    e->timing = enable != 0;
    e->ev_used = 0;
    e->ev_class.clear();
    return B200X_OK;
}

extern "C" int b200x_engine_get_timing(b200x_engine* e, double* ms_per_class, int64_t* launches_per_class) {
    B200X_REQUIRE(e && ms_per_class && launches_per_class, "get_timing: bad argument");
    B200X_CUDA_TRY(cudaStreamSynchronize(e->stream));
    for (int i = 0; i < KC_COUNT; ++i) { ms_per_class[i] = 0.0; launches_per_class[i] = 0; }
    for (size_t i = 0; i < e->ev_class.size(); ++i) {
        float ms = 0.f;
        B200X_CUDA_TRY(cudaEventElapsedTime(&ms, e->ev_pool[2 * i], e->ev_pool[2 * i + 1]));
        ms_per_class[e->ev_class[i]] += ms;
        launches_per_class[e->ev_class[i]] += 1;
    }
    e->ev_used = 0;
    e->ev_class.clear();
    return B200X_OK;
}
