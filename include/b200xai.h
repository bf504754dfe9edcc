/* b200xai.h - C ABI of the B200-native perturbation-explainability engine (libb200xai.so).
 *
 * Drop-in boundary for the hot path of Michal2711/Audio-Deepfake-Explainability.  The reference has no FFI of
 * its own (it is pure Python); the "plugin interface" it exposes for this path is the duck-typed predictor
 *      predict(wave: np.ndarray, sr: int) -> float                       (src/sonics_api.py:259-271)
 * and the two serial loops that call it once per perturbed copy
 *      SpectrogramExplainability._compute_occlusion_map                  (src/spectrogram_explainability.py:589-720)
 *      FrequencyBandPerturbation._compute_component_importance           (src/dsp_band_ops.py:529-666)
 *      predict_fn_unified (AudioLIME stem recombinations)                (src/lime_explainer.py:283-301)
 * Each engine-level entry point below replaces one of those call sites with a batched device sweep; the
 * kernel-level entry points expose every CUDA kernel separately so parity tests can check each stage.
 * INTEGRATION.md shows the ctypes binding a maintainer of the reference would add.
 *
 * Conventions: plain pointers and sizes only; every function returns B200X_OK (0) or an error code and
 * b200x_last_error() holds the message (thread-local).  There is NO CPU fallback and no silent 0.0 result
 * (contrast src/spectrogram_explainability.py:357-362): a failure is always reported.
 * Pointers prefixed d_ are device pointers; "stream" is a cudaStream_t passed as void* (NULL = default stream).
 */
#ifndef B200XAI_H
#define B200XAI_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200X_OK 0
#define B200X_ERR_INVALID 1
#define B200X_ERR_CUDA 2
#define B200X_ERR_STATE 3

#define B200X_GEMM_OUT_BF16 0      /* out = bf16(act(acc + bias))                                   */
#define B200X_GEMM_OUT_F32_RESID 1 /* out = resid + acc + bias              (fp32, in place allowed) */
#define B200X_GEMM_OUT_F32_TOKEN 2 /* out = act(acc + bias) + pe[row % group_in]  (fp32, row remap)  */

#define B200X_MASK_NONE 0
#define B200X_MASK_OCCLUDE 1   /* S[f0:f1, t0:t1] = value          (spectrogram_explainability.py:670-671) */
#define B200X_MASK_BAND_GAIN 2 /* S *= gain[f]                     (dsp_band_ops.py:576-579)               */
#define B200X_MASK_KEEP_ONLY 3 /* zeros except S[f0:f1, t0:t1]     (spectrogram_explainability.py:472-475) */
#define B200X_MASK_RANDOM_KEEP 4 /* S *= Bernoulli(p) bit per cell (RISE, spectrogram_explainability.py:768-769); the row of
                                  * d_windows is (seed, mask index, floor(p 2^32) as int32 bits, 0); the bit of cell
                                  * (frame t, bin k) is hash(seed, mask, t * 1025 + k) < threshold (common.h: rise_keep) */

const char* b200x_last_error(void);
int b200x_version(void);
/* select the CUDA device used by subsequent calls from this thread (one process per GPU: pass LOCAL_RANK) */
int b200x_set_device(int device);
int b200x_device_count(int* count);

/* ------------------------------------------------------------------------------------------------------------
 * Kernel-level entry points (device pointers)
 * ---------------------------------------------------------------------------------------------------------- */

/* librosa.stft(y, 2048, 512, 2048, 'hann', center=True) [reflect_pad = 0: zero padding] or the classifier's
 * torch.stft framing [reflect_pad = 1].  d_spec: complex64 interleaved, frame-major [1 + n/hop][spec_stride].
 * replaces: librosa.stft call, src/spectrogram_explainability.py:379-386, src/dsp_band_ops.py:394-401 */
int b200x_stft(const float* d_wave, int64_t n_samples, int n_fft, int hop, int reflect_pad, void* d_spec,
               int spec_stride, void* stream);

/* librosa.stft (zero padding) of `copies` equal-length waves in one launch: wave c at d_waves + c * wave_stride floats,
 * spectrum c at d_spec + c * spec_copy_stride complex values (Griffin-Lim: rebuilt = stft(istft(.)), librosa.griffinlim) */
int b200x_stft_batch(const float* d_waves, int64_t n_samples, int64_t wave_stride, int copies, void* d_spec, int spec_stride,
                     int64_t spec_copy_stride, void* stream);

/* Batched librosa.istft of `copies` perturbed versions of one spectrogram; the perturbation is applied in the
 * load stage (mode = B200X_MASK_*), d_windows int32 [copies][4] = t0,t1,f0,f1, d_gains float [copies][1025].
 * Writes hop*(n_frames-1) samples per copy at d_y + copy*y_stride; d_sumsq (optional, pre-zeroed double[copies])
 * receives sum(y^2) for match_rms.  d_frame_range (optional, int32 [copies][2] = [ma, mb) from b200x_frame_ranges,
 * max_range_frames = max(mb - ma)): only the samples that classifier frames [ma, mb) read are synthesised - by
 * linearity of the iSTFT every other sample equals the unperturbed track's.
 * replaces: src/spectrogram_explainability.py:670-680, dsp_band_ops.py:578-580 */
int b200x_istft_masked(const void* d_spec, int spec_stride, int n_frames, int copies, int mode,
                       const int32_t* d_windows, float occlusion_value, const float* d_gains, float* d_y,
                       int64_t y_stride, double* d_sumsq, const int32_t* d_frame_range, int max_range_frames,
                       void* stream);

/* Same, for copies that belong to SEVERAL tracks (FBP over a batch of tracks): copy c reads the spectrogram that starts
 * (c / copies_per_track) * track_stride complex values after d_spec. */
int b200x_istft_masked_tracks(const void* d_spec, int spec_stride, int n_frames, int copies, int copies_per_track,
                              int64_t track_stride, int mode, const int32_t* d_windows, float occlusion_value,
                              const float* d_gains, float* d_y, int64_t y_stride, double* d_sumsq,
                              const int32_t* d_frame_range, int max_range_frames, void* stream);

/* classifier frames [ma, mb) = [t0 - 4, t1 + 4) whose input samples an occlusion window t0,t1,f0,f1 can change */
int b200x_frame_ranges(const int32_t* d_windows, int n, int n_frames, int32_t* d_ranges, void* stream);

/* Classifier front-end (sonics FeatureExtractor = torchaudio MelSpectrogram + AmplitudeToDB, third-party):
 * reflect-padded STFT -> power -> HTK mel -> 10 log10(max(., amin)).  d_db float [copies][db_frames][n_mels] (row
 * t - ma when d_frame_range is given, else row t); d_cta_max float [copies][ceil(span / b200x_mel_frames_per_cta())].
 * If d_sumsq != NULL the samples are scaled by ref_rms / sqrt(sumsq/rms_count + 1e-8) first (match_rms,
 * src/dsp_band_ops.py:228-233). */
int b200x_mel_frames_per_cta(void);
int b200x_mel_db(const float* d_y, int64_t y_stride, int64_t n_samples, int copies, int sample_rate, int n_mels,
                 double f_min, double f_max, double amin, const double* d_sumsq, double ref_rms, int64_t rms_count,
                 float* d_db, int db_frames, float* d_cta_max, const int32_t* d_frame_range, int max_range_frames,
                 void* stream);
/* Same with one reference RMS per copy (d_ref_rms_per_copy[c] < 0 leaves copy c unscaled); NULL = the scalar ref_rms. */
int b200x_mel_db_ref(const float* d_y, int64_t y_stride, int64_t n_samples, int copies, int sample_rate, int n_mels,
                     double f_min, double f_max, double amin, const double* d_sumsq, double ref_rms,
                     const double* d_ref_rms_per_copy, int64_t rms_count, float* d_db, int db_frames, float* d_cta_max,
                     const int32_t* d_frame_range, int max_range_frames, void* stream);
/* d_out[i * repeat + k] = sqrt(mean(wave_i^2) + 1e-8), float64: the match_rms reference level (src/dsp_band_ops.py:228-233) */
int b200x_wave_rms(const float* d_waves, int64_t n_samples, int64_t stride, int n_waves, int repeat, double* d_out, void* stream);


/* prefix / suffix maxima of a baseline dB spectrogram: premax[m] = max over frames < m, sufmax[m] = max over frames >= m
 * (m = 0..n_frames), used for the per-copy top_db floor when only frames [ma, mb) were recomputed */
int b200x_mel_base_maxima(const float* d_db_base, int n_frames, int n_mels, float* d_premax, float* d_sufmax, void* stream);

/* top_db clamp, (x-mean)/(std+eps), F.interpolate(bilinear) along time to out_t, bf16, in both tokenizer operand
 * layouts: d_img_t [copies][out_t][n_mels], d_img_f [copies][n_mels][ld_f] (d_img_f may be NULL: the transposed copy is
 * skipped).  d_partial: 32*copies double2 scratch.
 * With d_db_base (+ maxima + d_frame_range) frames outside [ma, mb) are read from the baseline spectrogram. */
int b200x_mel_normalize_resize(const float* d_db, int db_frames, const float* d_cta_max, int n_cta_max, int copies,
                               int n_frames, int n_mels, float top_db, int unbiased, float eps, int out_t,
                               const float* d_db_base, const float* d_base_premax, const float* d_base_sufmax,
                               const int32_t* d_frame_range, void* d_partial, float* d_floor, void* d_img_t,
                               void* d_img_f, int ld_f, void* stream);

/* Rational polyphase resampling by up / down with the odd-length zero-phase FIR d_h (float64, unit DC gain):
 * y[n] = up * sum_k h[n * down - k * up + (h_len - 1) / 2] x[k], n < n_out = ceil(n_in * up / down)  (the resampling step of
 * librosa.load(path, sr=...), src/spectrogram_explainability.py:601; librosa itself uses soxr_hq - see INTEGRATION.md) */
int b200x_resample_poly(const float* d_x, int64_t n_in, int up, int down, const double* d_h, int h_len, float* d_y,
                        int64_t n_out, void* stream);

/* y[b] = sum_i masks[b][i] * stems[i]   (src/lime_explainer.py:283-301 composition) */
int b200x_mix_stems(const float* d_stems, int64_t n_samples, int n_stems, const uint8_t* d_masks, int copies,
                    float* d_y, int64_t y_stride, void* stream);

/* ---- mel-domain explainer variant (spec_type: mel; src/spectrogram_explainability.py:367-377, 394-402).  The reference's
 * inverse (librosa mel_to_audio: L-BFGS-B NNLS + Griffin-Lim from UNSEEDED random phases) is not reproducible, so these
 * kernels implement the builder-defined arithmetic restated in oracle/mel.py.  Spectra are frame-major, 1028 elements a row.
 * d_basis: dense float [n_mels][1025] filterbank; d_bin_range int32 [n_mels][2] non-zero bins of each filter; d_bin_first
 * int32 [1025] / d_bin_w float [1025][2]: the (at most two, adjacent) filters each bin feeds; d_pinv_t float [n_mels][1025]. */
/* d_mel[t][i] = sum_k basis[i][k] |S[t][k]|^2   (librosa.feature.melspectrogram, power 2) */
int b200x_mel_power(const void* d_spec, int spec_stride, int n_frames, int n_mels, const float* d_basis, const int32_t* d_bin_range,
                    float* d_mel, void* stream);
/* per copy and frame t in d_frame_range[copy] = [fa, fb): B = mel[t] with the perturbation applied (mode B200X_MASK_NONE /
 * OCCLUDE / KEEP_ONLY with d_windows (t0,t1,mel0,mel1) / BAND_GAIN with d_gains [copies][n_mels]); X = max(0, pinv B), then
 * nnls_iter steps X <- max(0, X - step basis^T (basis X - B)); d_mag[copy][t][k] = sqrt(X[k])  (mel_to_stft, power 2) */
int b200x_mel_nnls(const float* d_mel, int n_frames, int n_mels, int copies, int mode, const int32_t* d_windows,
                   float occlusion_value, const float* d_gains, const float* d_basis, const int32_t* d_bin_range,
                   const int32_t* d_bin_first, const float* d_bin_w, const float* d_pinv_t, float step, int nnls_iter,
                   const int32_t* d_frame_range, int max_range_frames, float* d_mag, int64_t mag_copy_stride, void* stream);
/* Griffin-Lim: C = mag * exp(2 pi i u) with u = hash(seed, first_index + copy, cell) / 2^32 (the RISE hash), and the update
 * a = rebuilt - coef * tprev; C = mag * a / (|a| + tiny)   (librosa.griffinlim, momentum: coef = m / (1 + m); 0 first) */
int b200x_gl_init(const float* d_mag, int64_t mag_copy_stride, void* d_c, int64_t c_copy_stride, int copies, int n_frames,
                  uint32_t seed, int first_index, void* stream);
int b200x_gl_update(const void* d_rebuilt, const void* d_tprev, const float* d_mag, int64_t mag_copy_stride, void* d_c,
                    int64_t c_copy_stride, int copies, int n_frames, float coef, void* stream);

/* tcgen05 GEMM  C[M,N] = A[M,K] . W[N,K]^T, bf16 operands (row-major, K contiguous), fp32 accumulation in TMEM,
 * fused epilogue per B200X_GEMM_OUT_*.  block_n in {128,192,208,256}.  B200X_GEMM_OUT_F32_RESID accumulates in place
 * (d_resid must equal d_out; TMA reduce-add).  Replaces every nn.Linear / Conv1d contraction of the third-party
 * SpecTTTra forward. */
int b200x_gemm_bf16(const void* d_a, int lda, const void* d_w, int ldw, int M, int N, int K, int block_n, void* d_out,
                    int ldc, int out_mode, const float* d_bias, int act_gelu, const float* d_resid, const float* d_pe,
                    int group_in, int group_out, int group_off, int reverse, void* stream);

/* Token-mode GEMM (B200X_GEMM_OUT_F32_TOKEN with group_in = 128) whose A operand is M-MAJOR: d_img bf16 [batch][K][128], row tile
 * b = the 128 contiguous indices of batch element b; out_row = b * group_out + group_off + m; out = act(A . W^T + bias) + pe[m].
 * The spectral tokenizer (Conv1d over the mel axis of the [mel][time] image in the third-party encoder) reads the [time][mel]
 * image the temporal tokenizer uses through an M-major tcgen05 operand descriptor, so the transposed copy is never written. */
int b200x_gemm_tokens_mmajor(const void* d_img, int batch, int K, const void* d_w, int ldw, int N, float* d_out, int ldc,
                             const float* d_bias, int act_gelu, const float* d_pe, int group_out, int group_off, void* stream);

/* The A-stationary CTA-pair kernel that b200x_gemm_bf16 selects for wide bf16 outputs of a narrow K at large M (the QKV
 * projection, nn.Linear(384, 1152) of the third-party encoder block): a CTA keeps its 128 x K row tile of A in shared memory and
 * sweeps every 192-column tile of W over it, so A crosses L2 -> SM once instead of once per column tile.  Same MMAs in the same
 * order as the pair kernel: bit-identical results.  bf16 output, optional bias / GELU; K a multiple of 64 up to 384; block_n 192
 * or 208. */
int b200x_gemm_bf16_astationary(const void* d_a, int lda, const void* d_w, int ldw, int M, int N, int K, int block_n, void* d_out,
                                int ldc, const float* d_bias, int act_gelu, int reverse, void* stream);

/* x += A . W^T + bias on the fp32 residual stream, then h = bf16(LayerNorm(x) * gamma + beta): a residual projection of a
 * pre-norm encoder block (attention proj, fc2) fused with the LayerNorm that FOLLOWS it (x + sublayer(x) then nn.LayerNorm in
 * the third-party encoder block reached through src/sonics_api.py:259-271).  Each CTA normalises its rows right after its
 * reduce-adds completed, from L2: the residual stream is not read back from HBM.  Same arithmetic as b200x_gemm_bf16
 * (B200X_GEMM_OUT_F32_RESID) followed by b200x_layernorm, bit for bit.  N = LayerNorm width (multiple of 128, <= 384);
 * d_a bf16 [M][lda], d_w bf16 [N][ldw], d_x fp32 [M][ldx] (in/out), d_h bf16 [M][ldh] (out). */
int b200x_gemm_resid_ln_bf16(const void* d_a, int lda, const void* d_w, int ldw, int M, int N, int K, float* d_x, int ldx,
                             const float* d_bias, const float* d_gamma, const float* d_beta, float eps, void* d_h, int ldh,
                             int reverse, void* stream);

/* fused softmax(Q K^T / sqrt(d)) V for d_qkv bf16 [copies*tokens][3*heads*64] = [q|k|v]; d_out bf16
 * [copies*tokens][heads*64].  (F.scaled_dot_product_attention inside the third-party encoder) */
int b200x_attention(const void* d_qkv, void* d_out, int copies, int tokens, int heads, int head_dim, int reverse,
                    void* stream);

/* `reverse` (GEMM CTA-pair kernel, attention, LayerNorm): != 0 walks tiles, (copy, head) blocks and rows from the end.
 * Results do not depend on it; the engine alternates it between consecutive kernels of the forward so that each one starts
 * on the part of its input that the previous kernel wrote last and that is still in L2.  It is a per-launch argument:
 * the library keeps no process-wide mutable state (engines on several GPUs / host threads do not interact). */

/* LayerNorm over dim (fp32 in); rows with (row % group) >= split use (gamma2, beta2) when group > 0.
 * Exactly one of d_out_bf16 / d_out_f32 (may alias d_x) is non-NULL. */
int b200x_layernorm(const float* d_x, int rows, int dim, const float* d_gamma, const float* d_beta,
                    const float* d_gamma2, const float* d_beta2, int group, int split, float eps, void* d_out_bf16,
                    float* d_out_f32, int reverse, void* stream);

/* final LayerNorm (optional) + token mean + Linear(dim,1) + sigmoid.  d_partial: copies*b200x_head_slices() floats. */
int b200x_head_slices(void);
int b200x_head(const float* d_x, int copies, int tokens, int dim, const float* d_gamma, const float* d_beta, float eps,
               int use_norm, const float* d_w, float bias, float* d_partial, float* d_logit, float* d_prob,
               void* stream);

/* delta[i] = (double)baseline - (double)prob[i]      (importance = baseline_pred - occluded_pred, :684) */
int b200x_delta(const float* d_prob, float baseline, int n, double* d_delta, void* stream);
/* same with the baseline read from device memory (no host round trip between the sweep and the reductions) */
int b200x_delta_dev(const float* d_prob, const float* d_baseline, int n, double* d_delta, void* stream);

/* importance map: map[f0:f1,t0:t1] += delta; count += 1; map /= count + 1e-8  (float64 [n_freq][n_time], window order;
 * src/spectrogram_explainability.py:695-696, 707) */
int b200x_saliency_reduce(const int32_t* d_windows, const double* d_delta, int n_windows, int n_freq, int n_time,
                          double* d_map, void* stream);

/* RISE: map[f][t] = sum_i keep_i(f, t) * pred[i] / (n_masks * p + 1e-8), float64 in mask order; the keep bits are re-derived
 * from the hash of B200X_MASK_RANDOM_KEEP (src/spectrogram_explainability.py:783, :798; min-max scaling :801 is the caller's) */
int b200x_rise_map(const double* d_pred, int n_masks, uint32_t seed, double keep_probability, int n_freq, int n_time,
                   double* d_map, void* stream);

/* FBP rows: map[rows[b][0]:rows[b][1], :] += delta[b]  (src/dsp_band_ops.py:652-653) */
int b200x_band_map(const int32_t* d_band_rows, const double* d_delta, int n_bands, int n_freq, int n_time,
                   double* d_map, void* stream);

/* stable ordering (Python sorted() semantics): mode 0 |v| desc, 1 |v| asc, 2 v desc, 3 v asc
 * (src/spectrogram_explainability.py:428-434, 566-571) */
int b200x_rank(const double* d_values, int n, int mode, int32_t* d_order, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Engine-level entry points (host or device buffers; an engine belongs to the CUDA device that was current at creation and
 * makes it current in every call; engines on different GPUs may coexist in one process; not re-entrant per engine)
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct b200x_engine b200x_engine;

typedef struct b200x_model_config {
    int32_t sample_rate, n_fft, hop_length, n_mels;
    double f_min, f_max, top_db, amin;
    float norm_eps;
    int32_t std_unbiased;
    int32_t input_spec_dim, input_temp_dim, t_clip, f_clip;
    int32_t embed_dim, num_heads, num_layers, mlp_hidden;
    int32_t pre_norm, pe_learnable, qkv_bias, final_norm;
    float tokenizer_ln_eps, block_ln_eps;
} b200x_model_config;

/* copies_per_chunk (1..1024): perturbed copies processed per pass; large chunks amortise per-launch costs (~22 MB of
 * device workspace per copy at 120 s / 16 kHz). */
int b200x_engine_create(const b200x_model_config* cfg, int copies_per_chunk, int64_t max_samples, b200x_engine** out);
void b200x_engine_destroy(b200x_engine* e);

/* sonics state-dict entry by name (float32 host data), then finalize (packs bf16 operands; all names required). */
int b200x_engine_set_param(b200x_engine* e, const char* name, const float* data, int64_t numel);
int b200x_engine_finalize(b200x_engine* e);

/* LocalSonnics.predict for `count` equal-length waves: fake-probability per wave (src/sonics_api.py:259-271). */
int b200x_engine_predict(b200x_engine* e, const float* waves, int64_t n_samples, int count, int on_device, float* prob,
                         float* logit /* nullable */);

/* Load one track: uploads the wave, computes the explainer STFT (librosa semantics) and keeps both resident. */
/* fake-probability (and logit, if not NULL) of the track loaded with set_track, read from its device-resident samples:
 * the baseline prediction of src/spectrogram_explainability.py:605 / dsp_band_ops.py:544 without a second upload */
int b200x_engine_predict_track(b200x_engine* e, float* prob, float* logit);
int b200x_engine_set_track(b200x_engine* e, const float* wave, int64_t n_samples, int on_device);
int b200x_engine_track_shape(b200x_engine* e, int32_t* n_freq, int32_t* n_time);
/* complex64 [n_freq][n_time] interleaved, the layout of librosa.stft / OcclusionResult.S */
int b200x_engine_get_spectrogram(b200x_engine* e, float* spec_host);

/* The occlusion hot loop (src/spectrogram_explainability.py:665-703) for `n` windows (int32 [n][4] = t0,t1,f0,f1):
 * prob[i] = predict(istft(S with window i set to occlusion_value)).  Buffers are host (on_device=0) or device. */
int b200x_engine_occlusion_sweep(b200x_engine* e, const int32_t* windows, int n, float occlusion_value, int on_device,
                                 float* prob);
/* occlusion_sweep plus the baseline prediction of the track itself (base_prob[0] == predict_track, bit for bit) in the same
 * device pass: the track rides as one more copy of the last chunk when there is room for it. */
int b200x_engine_occlusion_sweep_base(b200x_engine* e, const int32_t* windows, int n, float occlusion_value, int on_device,
                                      float* prob, float* base_prob);
/* The FBP hot loop (src/dsp_band_ops.py:573-586) for `n` band gains (float [n][n_freq] = keep + att*(1-keep)). */
int b200x_engine_fbp_sweep(b200x_engine* e, const float* gains, int n, int normalize_loudness, int on_device,
                           float* prob);
/* FBP over a batch of equal-length tracks (host buffers): waves float [n_tracks][n_samples], gains float [n_bands][1025]
 * shared by all tracks; base_prob float [n_tracks] = predict(track), prob float [n_tracks][n_bands].  The band copies of as
 * many tracks as fit one chunk share one iSTFT launch and one classifier forward; per-copy results are identical to
 * set_track + predict_track + fbp_sweep track by track.  The last track is left as the engine's current track. */
int b200x_engine_fbp_sweep_tracks(b200x_engine* e, const float* waves, int n_tracks, int64_t n_samples, const float* gains,
                                  int n_bands, int normalize_loudness, float* base_prob, float* prob);
/* AudioLIME recombinations (src/lime_explainer.py:283-301): stems float [n_stems][n_samples], masks uint8 [n][n_stems]. */
int b200x_engine_stem_sweep(b200x_engine* e, const float* stems, int n_stems, int64_t n_samples, const uint8_t* masks,
                            int n, int on_device, float* prob);
/* Top-window audio reconstructed from the patch alone (src/spectrogram_explainability.py:472-483): row i of audio_host
 * (stride seg_stride floats) receives y_patch[t0*hop : min(t0*hop + max(1,(t1-t0)*hop), hop*(n_time-1))]; its length is
 * written to seg_len[i] (nullable). */
int b200x_engine_window_audio(b200x_engine* e, const int32_t* windows, int n, float* audio_host, int64_t seg_stride,
                              int64_t* seg_len);
/* Full occluded waveforms istft(S with window i := value): float [n][hop*(n_time-1)] host (lets a foreign duck-typed
 * predictor consume the GPU DSP stage; src/spectrogram_explainability.py:670-680). */
int b200x_engine_occluded_audio(b200x_engine* e, const int32_t* windows, int n, float occlusion_value, float* audio_host);
/* Perturbed audio of the FBP bands (for separated_bands WAVs, src/dsp_band_ops.py:608-639): float [n][hop*(n_time-1)]. */
int b200x_engine_band_audio(b200x_engine* e, const float* gains, int n, float* audio_host);

/* Track-loader front: b200x_resample_poly on host buffers (upload, one launch, read back) */
int b200x_engine_resample(b200x_engine* e, const float* x_host, int64_t n_in, int up, int down, const double* h_host, int h_len,
                          float* y_host, int64_t n_out);

/* Reductions on host buffers (map is float64 [n_freq][n_time]). */
int b200x_engine_saliency_map(b200x_engine* e, const int32_t* windows, const double* delta, int n, double* map_host);
/* RISE (src/spectrogram_explainability.py:722-806): probabilities of masks first_mask .. first_mask + n - 1 of the track set
 * with set_track (mask generated in the iSTFT load stage, never stored), their masked audio (tests), and the accumulated map */
int b200x_engine_rise_sweep(b200x_engine* e, int first_mask, int n, uint32_t seed, double keep_probability, int on_device,
                            float* prob);
int b200x_engine_rise_audio(b200x_engine* e, int first_mask, int n, uint32_t seed, double keep_probability, float* audio_host);
int b200x_engine_rise_map(b200x_engine* e, const double* pred, int n, uint32_t seed, double keep_probability, double* map_host);
int b200x_engine_band_map(b200x_engine* e, const int32_t* band_rows, const double* delta, int n, double* map_host);
/* saliency_map over an arbitrary [n_freq][n_time] grid (the mel variant's map has n_mels rows) */
int b200x_engine_saliency_map_shape(b200x_engine* e, const int32_t* windows, const double* delta, int n, int n_freq, int n_time,
                                    double* map_host);

/* Mel-domain variant.  set_mel_basis: dense filterbank float [n_mels][1025] (librosa.filters.mel, Slaney), its pseudo-inverse
 * float [1025][n_mels] and 1 / ||basis||_2^2 (host float64 linear algebra, mel_host.py).  mel_spectrogram: power mel of the
 * current track, float [n_mels][n_time] (librosa layout).  mel_sweep: for every perturbed copy i
 *     y_i = griffinlim(sqrt(nnls(basis, perturbed mel)), n_iter, phases of index first_index + i), prob[i] = predict(y_i)
 * (src/spectrogram_explainability.py:663-703 with spec_type == 'mel'); windows int32 [n][4] = t0,t1,mel0,mel1 for
 * B200X_MASK_OCCLUDE / KEEP_ONLY, gains float [n][n_mels] for BAND_GAIN (the builder's FBP-mel); prob and audio_host
 * (float [n][hop * (n_time - 1)]) are host buffers, either may be NULL. */
int b200x_engine_set_mel_basis(b200x_engine* e, int n_mels, const float* basis, const float* pinv, float step);
int b200x_engine_mel_spectrogram(b200x_engine* e, float* mel_host);
int b200x_engine_mel_sweep(b200x_engine* e, int mode, const int32_t* windows, const float* gains, int n, float occlusion_value,
                           int n_iter, int nnls_iter, uint32_t seed, int first_index, float momentum, float* prob,
                           float* audio_host);
int b200x_engine_rank(b200x_engine* e, const double* values, int n, int mode, int32_t* order_host);

/* Introspection for tests / profiling: device pointer of a named intermediate of the LAST processed chunk
 * ("y","db","img_t","img_f","x","h","qkv","att","hid","prob","S"), its byte size; number of kernel launches so far. */
int b200x_engine_debug_buffer(b200x_engine* e, const char* name, void** d_ptr, int64_t* bytes);
int b200x_engine_set_trace(b200x_engine* e, float* d_trace /* [layers+1][copies*tokens][dim] or NULL */);
int64_t b200x_engine_launch_count(b200x_engine* e);
/* per-kernel-class CUDA-event timing on the engine stream: classes 0 istft, 1 mel, 2 normalise/resize, 3 gemm,
 * 4 attention, 5 layernorm, 6 head, 7 other; get_timing returns the sums since set_timing / the last get (8 entries). */
int b200x_engine_set_timing(b200x_engine* e, int enable);
/* Alternate the traversal direction between consecutive kernels of the classifier forward (default on; see the `reverse`
 * launch argument above).  Results are unaffected. */
int b200x_engine_set_alternate(b200x_engine* e, int enable);
/* The classifier forward of a chunk is replayed from a CUDA graph once its shape has been seen twice (default on);
 * 0 = always launch kernel by kernel.  Results are identical either way. */
int b200x_engine_set_graphs(b200x_engine* e, int enable);
/* LayerNorm_1 of block l+1 as a tail of block l's fc2 GEMM (b200x_gemm_resid_ln_bf16, default on) or as a separate pass.
 * Results are bit-identical either way. */
int b200x_engine_set_fused_layernorm(b200x_engine* e, int enable);
int b200x_engine_get_timing(b200x_engine* e, double* ms_per_class, int64_t* launches_per_class);
void* b200x_engine_stream(b200x_engine* e);
int b200x_engine_synchronize(b200x_engine* e);

#ifdef __cplusplus
}
#endif
#endif /* B200XAI_H */
