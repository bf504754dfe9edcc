"""Engine-level parity: SpecTTTra forward, occlusion sweep, FBP sweep, stem sweep vs the CPU oracle
(tolerance from BASELINE.json north_star: 1e-3 absolute on per-window delta-prob and saliency values; window /
band indexing and the top-k set bit-exact)."""
import numpy as np
import pytest
import torch

from audio_deepfake_explainability_b200 import grid, synth
from audio_deepfake_explainability_b200.engine import Engine
from audio_deepfake_explainability_b200.weights import ALPHA_120S, random_state_dict
from oracle import dsp, loops, spectttra

pytestmark = pytest.mark.gpu
CFG = ALPHA_120S
TOL = 1e-3


@pytest.fixture(scope="module")
def sd():
    return random_state_dict(CFG, 0)


@pytest.fixture(scope="module")
def eng(sd):
    e = Engine(CFG, sd, copies_per_chunk=4, max_samples=30 * 16000)
    yield e
    e.close()


def test_forward_stages_match_oracle(eng, sd):
    y = synth.synth_track("REAL", 0, 16000, 20.0)
    M = CFG.num_tokens
    trace = torch.zeros(CFG.num_layers + 1, M, CFG.embed_dim, device="cuda")
    eng.set_trace(trace.data_ptr())
    try:
        prob, logit = eng.predict(y, return_logits=True)
    finally:
        eng.set_trace(None)
    t = torch.from_numpy(y).unsqueeze(0)
    img = spectttra.resize(dsp.mel_frontend(t, CFG), CFG)
    for mode, tol_tok, tol_layer, tol_logit in (("bf16", 6e-3, 3e-2, 3e-3), ("fp32", 3e-2, 1e-1, 1.5e-2)):
        tok = spectttra.tokenize(img, sd, CFG, mode)
        _, layers = spectttra.encoder(tok, sd, CFG, mode, return_all=True)
        e0 = (trace[0].cpu() - tok[0]).abs().max().item()
        assert e0 < tol_tok, f"{mode}: tokenizer max err {e0}"
        for l, ref in enumerate(layers):
            el = (trace[l + 1].cpu() - ref[0]).abs().max().item()
            assert el < tol_layer * max(1.0, ref.abs().max().item() / 4), f"{mode}: layer {l} max err {el}"
        ref_logit = spectttra.forward_logits(t, sd, CFG, mode).item()
        assert abs(float(logit) - ref_logit) < tol_logit, f"{mode}: logit {float(logit)} vs {ref_logit}"
        assert abs(float(prob) - 1.0 / (1.0 + np.exp(-ref_logit))) < TOL


def test_predict_batch_equals_single(eng):
    ys = np.stack([synth.synth_track(f, 1, 16000, 6.0) for f in ("REAL", "SUNO", "UDIO", "ElevenLabs", "SUNO_PRO")])
    pb = eng.predict(ys)
    ps = np.array([eng.predict(y) for y in ys])
    assert np.array_equal(pb, ps)                    # batching / chunking must not change a single bit
    assert len(set(np.round(pb, 6))) > 1


def test_occlusion_sweep_matches_oracle(eng, sd):
    y = synth.synth_track("UDIO", 0, 16000, 24.0)[: 16000 * 24 - 200]      # ragged length: iSTFT shorter than y
    eng.set_track(y)
    n_freq, n_time = eng.track_shape()
    assert (n_freq, n_time) == grid.stft_shape(len(y), 2048, 512)
    S = eng.spectrogram()
    S_ref = dsp.stft(y).numpy()
    assert np.abs(S - S_ref).max() < 2e-5 * np.abs(S_ref).max()
    wins = grid.occlusion_windows(n_freq, n_time, 256, 256, 20.0, 20.0)
    sel = np.concatenate([wins[:3], wins[-3:], wins[len(wins) // 2: len(wins) // 2 + 2]])
    base = float(eng.predict(y))
    prob = eng.occlusion_sweep(sel, 0.0)
    delta = np.float64(np.float32(base)) - prob.astype(np.float64)
    pred32 = spectttra.OraclePredictor(sd, CFG, "fp32")
    base_ref = pred32.predict(y, 16000)
    assert abs(base - base_ref) < TOL
    for i, (t0, t1, f0, f1) in enumerate(sel):
        S_occ = S_ref.copy()
        S_occ[f0:f1, t0:t1] = 0.0
        y_occ = dsp.istft(S_occ).numpy()
        y_occ = np.pad(y_occ, (0, len(y) - len(y_occ)))
        d_ref = base_ref - pred32.predict(y_occ, 16000)
        assert abs(delta[i] - d_ref) < TOL, f"window {i}: {delta[i]} vs {d_ref}"
    assert np.abs(delta).max() > 1e-4                 # the comparison is not vacuous
    # reductions through the engine: map identical to the reference loop given the same deltas
    m = eng.saliency_map(sel, delta)
    assert np.array_equal(m, loops.saliency_from_windows(sel, delta, n_freq, n_time))
    for mode, key, desc in ((0, abs, True), (1, abs, False), (2, float, True), (3, float, False)):
        assert eng.rank(delta, mode).tolist() == sorted(range(len(delta)), key=lambda i: key(delta[i]), reverse=desc)
    # baseline inside the sweep (dense path here would be sparse: 256-frame windows on a 750-frame track): same bits as separate calls
    prob_b, base_b = eng.occlusion_sweep(sel[:3], 0.0, with_baseline=True)          # 3 + 1 copies = one chunk of 4
    assert np.float32(base_b) == np.float32(base) and np.array_equal(prob_b, prob[:3])
    prob_b, base_b = eng.occlusion_sweep(sel, 0.0, with_baseline=True)              # 8 windows = 2 full chunks + the baseline alone
    assert np.float32(base_b) == np.float32(base) and np.array_equal(prob_b, prob)
    # chunking invariance (copies_per_chunk = 4, 8 windows = 2 chunks) and determinism
    assert np.array_equal(prob, eng.occlusion_sweep(sel, 0.0))
    assert np.array_equal(prob[5:6], eng.occlusion_sweep(sel[5:6], 0.0))
    # top-window audio (keep-only iSTFT)
    aud = eng.window_audio(sel[:2])
    for i in range(2):
        p = dict(t_start=int(sel[i][0]), t_end=int(sel[i][1]), f_start=int(sel[i][2]), f_end=int(sel[i][3]))
        ref = loops.window_audio(y, S_ref, p, 512, 2048, use_original_audio=False)
        assert aud[i].shape == ref.shape
        assert np.abs(aud[i] - ref).max() < 3e-6


@pytest.mark.parametrize("normalize", [False, True])
def test_fbp_sweep_matches_oracle(eng, sd, normalize):
    y = synth.synth_track("SUNO", 2, 16000, 16.0)
    eng.set_track(y)
    bands = grid.FREQUENCY_BAND_PRESETS["high_resolution"]
    sel = [bands[i] for i in (0, 3, 6, 8, 12)]         # incl. an empty band above Nyquist
    gains = grid.band_gain_table(sel, 16000, 2048, 0.25, "rel", 0.2, 5.0, 500.0, 200.0)
    base = float(eng.predict(y))
    prob = eng.fbp_sweep(gains, normalize)
    delta = np.float64(np.float32(base)) - prob.astype(np.float64)
    ref = loops.fbp_component(y, spectttra.OraclePredictor(sd, CFG, "fp32"), 16000, sel, 0.25, "rel", 200.0, 0.2, 5.0, 500.0,
                              normalize_loudness=normalize)
    ref_delta = np.array([b["importance"] for b in ref.batch_importances])
    assert np.abs(delta - ref_delta).max() < TOL, (delta, ref_delta)
    assert abs(delta[-1]) < TOL                         # empty band: keep == 1 everywhere
    rows = grid.band_bin_ranges(sel, 16000, 2048)
    assert np.array_equal(eng.band_map(rows, ref_delta), ref.importance_map)


def test_stem_sweep_matches_oracle(eng, sd):
    y, stems = synth.synth_track("REAL", 3, 16000, 8.0, with_stems=True)
    st = np.stack([stems[k] for k in sorted(stems)])
    masks = np.array([[1, 1, 1, 1], [1, 0, 0, 0], [0, 1, 1, 0], [0, 0, 0, 0]], np.uint8)
    prob = eng.stem_sweep(st, masks)
    ref = loops.stem_mask_probs(st, masks, spectttra.OraclePredictor(sd, CFG, "fp32"), 16000)
    assert np.abs(prob - ref[:, 1]).max() < TOL


def test_errors_are_loud(sd):
    e = Engine(CFG, sd, copies_per_chunk=2, max_samples=8 * 16000)
    try:
        with pytest.raises(RuntimeError):
            e.occlusion_sweep(np.array([[0, 1, 0, 1]], np.int32))            # no track loaded
        e.set_track(synth.synth_track("REAL", 0, 16000, 4.0))
        with pytest.raises(RuntimeError):
            e.occlusion_sweep(np.array([[0, 99999, 0, 1]], np.int32))        # window outside the spectrogram
        with pytest.raises(RuntimeError):
            e.predict(np.zeros(16 * 16000, np.float32))                      # longer than max_samples
        assert e.occlusion_sweep(np.zeros((0, 4), np.int32)).shape == (0,)   # empty sweep is a no-op
    finally:
        e.close()
    bad = dict(sd)
    bad.pop("classifier.bias")
    with pytest.raises(RuntimeError):
        Engine(CFG, bad, copies_per_chunk=2, max_samples=8 * 16000)


def test_graph_replay_is_bit_identical_to_eager_launches(eng):
    """The CUDA-graph replay of a chunk (third and later sweeps of a shape) must reproduce the eager launches exactly."""
    y = synth.synth_track("SUNO", 3, 16000, 6.0)
    eng.set_track(y)
    n_freq, n_time = eng.track_shape()
    wins = grid.occlusion_windows(n_freq, n_time, 64, 32, 25.0, 12.5)[:11]
    eng.set_graphs(False)
    eager = eng.occlusion_sweep(wins, 0.0)
    eng.set_graphs(True)
    runs = [eng.occlusion_sweep(wins, 0.0) for _ in range(3)]      # eager warm-up, capture + launch, replay
    for r in runs:
        assert np.array_equal(r, eager)
    assert np.array_equal(eng.predict(np.stack([y, y[::-1].copy()])), np.concatenate([eng.predict(y)[None], eng.predict(y[::-1].copy())[None]]))


def test_layernorm_tail_is_bit_identical_to_the_separate_pass(eng):
    """fc2 + LayerNorm tail (b200x_gemm_resid_ln_bf16, the default) against residual GEMM followed by the LayerNorm pass: the
    same bits through twelve blocks, eager and from the replayed graph, for a ragged chunk (11 copies in chunks of 4)."""
    y = synth.synth_track("UDIO", 5, 16000, 6.0)
    eng.set_track(y)
    n_freq, n_time = eng.track_shape()
    wins = grid.occlusion_windows(n_freq, n_time, 64, 32, 25.0, 12.5)[:11]
    try:
        eng.set_fused_layernorm(False)
        separate = [eng.occlusion_sweep(wins, 0.0) for _ in range(3)]
        p_sep = eng.predict(y)
        eng.set_fused_layernorm(True)
        fused = [eng.occlusion_sweep(wins, 0.0) for _ in range(3)]
        p_fused = eng.predict(y)
    finally:
        eng.set_fused_layernorm(True)
    for a in separate + fused:
        assert np.array_equal(a, separate[0])
    assert p_sep == p_fused


@pytest.mark.parametrize("n_mels,embed_dim,heads", [(64, 384, 6), (96, 256, 4)])
def test_other_classifier_geometries_match_oracle(n_mels, embed_dim, heads):
    """Classifier configurations other than the alpha-120s default: n_mels != 128 takes the K-major spectral tokenizer with the
    transposed image copy (the 128-mel default reads one image through an M-major operand), embed_dim 256 takes the LayerNorm
    paths with two vectors per lane and the generic GEMM tile choices.  Tokenizer, every block and the logit against the oracle."""
    import dataclasses

    cfg = dataclasses.replace(CFG, n_mels=n_mels, input_spec_dim=n_mels, embed_dim=embed_dim, num_heads=heads, num_layers=3)
    sd2 = random_state_dict(cfg, 3)
    e = Engine(cfg, sd2, copies_per_chunk=3, max_samples=12 * 16000)
    try:
        ys = np.stack([synth.synth_track("SUNO", 1, 16000, 10.0), synth.synth_track("REAL", 2, 16000, 10.0)])
        M = cfg.num_tokens
        trace = torch.zeros(cfg.num_layers + 1, 2 * M, cfg.embed_dim, device="cuda")
        e.set_trace(trace.data_ptr())
        try:
            prob, logit = e.predict(ys, return_logits=True)
        finally:
            e.set_trace(None)
        t = torch.from_numpy(ys)
        img = spectttra.resize(dsp.mel_frontend(t, cfg), cfg)
        tok = spectttra.tokenize(img, sd2, cfg, "bf16")
        _, layers = spectttra.encoder(tok, sd2, cfg, "bf16", return_all=True)
        tr = trace.cpu().reshape(cfg.num_layers + 1, 2, M, cfg.embed_dim)
        assert (tr[0] - tok).abs().max().item() < 6e-3
        for l, ref in enumerate(layers):
            assert (tr[l + 1] - ref).abs().max().item() < 3e-2 * max(1.0, ref.abs().max().item() / 4), l
        ref_logit = spectttra.forward_logits(t, sd2, cfg, "fp32").numpy().reshape(-1)
        assert np.abs(np.asarray(logit).reshape(-1) - ref_logit).max() < 1.5e-2
        assert np.abs(np.asarray(prob).reshape(-1) - 1.0 / (1.0 + np.exp(-ref_logit))).max() < TOL
        assert np.array_equal(np.asarray(prob).reshape(-1), np.array([e.predict(y) for y in ys]).reshape(-1))
    finally:
        e.close()
