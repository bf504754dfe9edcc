"""The drop-in explainer classes end to end on a GPU: same entry points, return tuples and on-disk layout as the
reference (src/spectrogram_explainability.py:413-587, 589-720, 808-916; src/dsp_band_ops.py:516-527, 529-666), checked
against the CPU oracle loops on the same synthetic track and the same random-init weights."""
import json
from pathlib import Path

import numpy as np
import pytest

from audio_deepfake_explainability_b200 import grid, synth
from audio_deepfake_explainability_b200.audio_io import load_audio, write_wav
from audio_deepfake_explainability_b200.dsp_band_ops import FBDResult, FrequencyBandPerturbation
from audio_deepfake_explainability_b200.sonics_api import B200Predictor
from audio_deepfake_explainability_b200.spectrogram_explainability import OcclusionResult, SpectrogramExplainability
from audio_deepfake_explainability_b200.weights import ALPHA_120S, random_state_dict
from oracle import loops, spectttra                      # checker only

pytestmark = pytest.mark.gpu
TOL = 1e-3                                               # north_star: per-window delta-prob and saliency within 1e-3 absolute
SR = 16000


@pytest.fixture(scope="module")
def predictor():
    p = B200Predictor.random_init(seed=0, copies_per_chunk=8, max_samples=SR * 12)
    yield p
    p.close()


@pytest.fixture(scope="module")
def oracle_predictor():
    return spectttra.OraclePredictor(random_state_dict(ALPHA_120S, 0), ALPHA_120S, "fp32")


@pytest.fixture(scope="module")
def track(tmp_path_factory):
    root = tmp_path_factory.mktemp("data")
    (root / "UDIO").mkdir()
    y = synth.synth_track("UDIO", 3, SR, 10.0)
    path = root / "UDIO" / "3_UDIO.wav"
    write_wav(path, y, SR)
    y_back, sr = load_audio(str(path), sr=SR, duration=120, mono=True)
    assert sr == SR and len(y_back) == len(y)
    return root, path, y_back


def test_occlusion_explainer_matches_oracle_loop(predictor, oracle_predictor, track, tmp_path):
    root, path, y = track
    ex = SpectrogramExplainability(predictor, sr=SR, duration=120, spec_type="stft", method="occlusion", top_n_windows=3,
                                   use_original_audio=False, patch_time_frames=128, stride_time_frames=64,
                                   patch_freq_percent=25.0, stride_freq_percent=12.5)
    res = ex._compute_occlusion_map(str(path), occlusion_value=0.0, baseline_threshold=0.0, verbose=False)
    assert isinstance(res, OcclusionResult)
    ref = loops.occlusion_map(y, oracle_predictor, SR, patch_time_frames=128, stride_time_frames=64, patch_freq_percent=25.0,
                              stride_freq_percent=12.5, baseline_threshold=0.0)
    # window indexing bit-exact, delta-prob and saliency within tolerance
    assert [(p["t_start"], p["t_end"], p["f_start"], p["f_end"]) for p in res.patch_importances] == \
           [(p["t_start"], p["t_end"], p["f_start"], p["f_end"]) for p in ref.patch_importances]
    got = np.array([p["importance"] for p in res.patch_importances])
    want = np.array([p["importance"] for p in ref.patch_importances])
    assert abs(res.baseline_pred - ref.baseline_pred) < TOL
    assert np.abs(got - want).max() < TOL
    assert np.abs(want).max() > 1e-4                      # not vacuous
    assert res.importance_map.shape == ref.importance_map.shape == res.S.shape
    assert np.abs(res.importance_map - ref.importance_map).max() < TOL
    assert np.abs(res.S - ref.S).max() < 2e-5 * np.abs(ref.S).max()
    # given the SAME importances, the map and the four top-k groups are bit-exact with the reference loop
    w = np.array([[p["t_start"], p["t_end"], p["f_start"], p["f_end"]] for p in res.patch_importances], np.int32)
    assert np.array_equal(res.importance_map, loops.saliency_from_windows(w, got, *res.S.shape))
    groups = ex.top_window_groups(res.patch_importances, 3, "3_UDIO")
    assert groups == loops.top_window_groups(res.patch_importances, 3, "3_UDIO", 512, SR)


def test_occlusion_process_audio_file_writes_reference_layout(predictor, track, tmp_path):
    root, path, y = track
    ex = SpectrogramExplainability(predictor, sr=SR, duration=120, spec_type="stft", method="occlusion", top_n_windows=2,
                                   use_original_audio=False, patch_time_frames=128, stride_time_frames=128,
                                   patch_freq_percent=50.0, stride_freq_percent=50.0, checkpoint_dir=tmp_path / "ckpt")
    out = tmp_path / "saliency_maps"
    summary = ex.process_audio_file(str(path), out, baseline_threshold=0.0, folder_name="UDIO")
    assert summary["file_name"] == "3_UDIO" and summary["folder"] == "UDIO" and summary["method"] == "occlusion"
    tdir = out / "UDIO" / "3_UDIO" / "top_windows"
    for g in ("all", "best", "worst", "most_influential"):
        meta = json.loads((tdir / g / f"3_UDIO__{g}_occlusion_patches_from_list.json").read_text())
        assert meta["file_name"] == "3_UDIO" and meta["group"] == g and meta["top_n"] == len(meta["windows"])
        for k, m in enumerate(meta["windows"], 1):
            assert m["rank"] == k and m["type"] in ("POSITIVE", "NEGATIVE", "NEUTRAL")
            assert set(m) == {"rank", "t_start", "t_end", "f_start", "f_end", "start_time_sec", "end_time_sec", "importance",
                              "abs_importance", "type"}
        wavs = sorted((tdir / g).glob("*.wav"))
        if g == "all":
            assert not wavs                                # JSON only (:566-571)
        else:
            assert len(wavs) == len(meta["windows"])
            m = meta["windows"][0]
            name = (f"3_UDIO__{g}1_patch_{m['type']}_{m['abs_importance']:.3f}_t{m['t_start']}-{m['t_end']}"
                    f"_f{m['f_start']}-{m['f_end']}.wav")
            assert (tdir / g / name).exists()
            seg, _ = load_audio(str(tdir / g / name), sr=None, mono=True)
            assert len(seg) == (m["t_end"] - m["t_start"]) * 512
    assert np.load(out / "UDIO" / "3_UDIO" / "saliency_3_UDIO.npy").shape == (1025, 1 + len(y) // 512)
    # resume: the checkpoint marks the file and a second call skips it (:97-135, :815-818)
    assert ex.process_audio_file(str(path), out, baseline_threshold=0.0, folder_name="UDIO") is None
    # below the baseline threshold the file is skipped but marked processed (:609-619, :848-851)
    ex2 = SpectrogramExplainability(predictor, sr=SR, spec_type="stft", method="occlusion", patch_time_frames=128,
                                    stride_time_frames=128, patch_freq_percent=50.0, stride_freq_percent=50.0)
    r = ex2._compute_occlusion_map(str(path), baseline_threshold=1.1, verbose=False)
    assert r.importance_map is None and r.patch_importances is None


def test_unsupported_variants_fail_loudly(predictor):
    with pytest.raises(TypeError):
        SpectrogramExplainability(object(), sr=SR)
    ex = SpectrogramExplainability(predictor, sr=SR, spec_type="mel", method="rise")       # mel is built for occlusion only
    with pytest.raises(NotImplementedError):
        ex.rise_map_from_wave(np.zeros(SR, np.float32))
    ex = SpectrogramExplainability(predictor, sr=SR, spec_type="mel", method="occlusion", fmax=4000)
    with pytest.raises(NotImplementedError):                                                # forward / inverse filterbanks would differ (:375 vs :395)
        ex.occlusion_map_from_wave(np.zeros(SR, np.float32))
    ex = SpectrogramExplainability(predictor, sr=SR, spec_type="stft", method="rise")
    with pytest.raises(NotImplementedError):
        ex.occlusion_map_from_wave(np.zeros(SR, np.float32))
    ex = SpectrogramExplainability(predictor, sr=SR, spec_type="stft", method="occlusion", n_fft=1024)
    with pytest.raises(NotImplementedError):                                                # kernels are built for 2048 / 512 / 2048
        ex.occlusion_map_from_wave(np.zeros(SR, np.float32))
    with pytest.raises(ValueError):
        SpectrogramExplainability(predictor, sr=SR, spec_type="cqt")
    with pytest.raises(ValueError):
        FrequencyBandPerturbation(predictor, spec_type="cqt", sr=SR)


@pytest.mark.parametrize("normalize", [False, True])
def test_fbp_explainer_matches_oracle_loop(predictor, oracle_predictor, track, tmp_path, normalize):
    root, path, y = track
    kw = dict(preset="high_resolution", attenuation=0.25, transition_mode="rel", transition_rel=0.2, transition_min_hz=5.0,
              transition_max_hz=500.0)
    fbp = FrequencyBandPerturbation(predictor, sr=SR, normalize_loudness=normalize, **kw)
    res = fbp._compute_component_importance(y, "mixture", str(path))
    assert isinstance(res, FBDResult)
    ref = loops.fbp_component(y, oracle_predictor, SR, bands=fbp.bands, attenuation=0.25, transition_mode="rel",
                              transition_rel=0.2, transition_min_hz=5.0, transition_max_hz=500.0, normalize_loudness=normalize)
    assert [(b["component"], b["low"], b["high"]) for b in res.batch_importances] == \
           [(b["component"], b["low"], b["high"]) for b in ref.batch_importances]
    got = np.array([b["importance"] for b in res.batch_importances])
    want = np.array([b["importance"] for b in ref.batch_importances])
    assert abs(res.baseline_pred - ref.baseline_pred) < TOL
    assert np.abs(got - want).max() < TOL
    assert np.abs(res.importance_map - ref.importance_map).max() < 2 * TOL      # the shared 250 Hz bin sums two bands
    # band -> bin rows are bit-exact: the map's support equals the reference's (hard inclusive edges, empty bands above Nyquist)
    rows = grid.band_bin_ranges(fbp.bands, SR, 2048)
    covered = np.zeros(1025, bool)
    for lo, hi in rows:                                     # (first bin, last bin exclusive); an empty band is (0, 0)
        covered[lo:hi] = True
    assert np.array_equal(np.abs(ref.importance_map).sum(1) > 0, covered & (np.abs(ref.importance_map).sum(1) > 0))
    assert np.all(res.importance_map[~covered] == 0)
    out = tmp_path / "bands"
    summary = fbp.process_audio_file(str(path), out, folder_name="UDIO")
    meta = json.loads((out / "UDIO" / "3_UDIO" / "mixture" / "3_UDIO_bands_metadata.json").read_text())
    assert meta["file_name"] == "3_UDIO" and len(meta["bands"]) == 13
    assert set(meta["bands"][0]) == {"component", "low", "high", "importance", "abs_importance", "type"}
    assert set(summary["components"]) == {"mixture"}


def test_fbp_band_audio_layout(predictor, track, tmp_path):
    root, path, y = track
    fbp = FrequencyBandPerturbation(predictor, sr=SR, preset="default", attenuation=0.25, transition_mode="rel", transition_rel=0.2,
                                    transition_min_hz=5.0, transition_max_hz=500.0, normalize_loudness=False,
                                    save_perturbed_audio_only=True)
    assert fbp.process_audio_file(str(path), tmp_path / "bands", folder_name="UDIO") is None      # audio-save modes return None (:655-657)
    wavs = sorted((tmp_path / "bands" / "UDIO" / "3_UDIO" / "mixture" / "separated_bands" / "freq_batches").glob("*.wav"))
    assert len(wavs) == len(fbp.bands)
    assert all(w.name.startswith("3_UDIO__mixture__") and "Hz_" in w.name for w in wavs)


def test_stem_mask_sweep_matches_oracle(predictor, oracle_predictor):
    stems = np.stack(list(synth.synth_stems("REAL", 1, SR, 6.0).values())[:4]).astype(np.float32)
    masks = np.random.RandomState(0).randint(0, 2, 12 * 4).reshape(12, 4).astype(np.uint8)
    masks[0] = 1
    got = predictor.stem_mask_sweep(stems, masks)
    want = loops.stem_mask_probs(stems, masks, oracle_predictor, SR)
    assert got.shape == want.shape == (12, 2)
    assert np.abs(got - want).max() < TOL
    assert np.allclose(got.sum(1), 1.0)


def test_lime_explain_stems_matches_oracle(predictor, oracle_predictor):
    from audio_deepfake_explainability_b200 import lime_explainer as le
    stems = np.stack(list(synth.synth_stems("SUNO_PRO", 2, SR, 6.0).values())[:4]).astype(np.float32)
    exp = le.explain_stems(stems, predictor, num_samples=60, random_state=0)
    full = le.explain_stems(stems, predictor, num_samples=60, random_state=0, deduplicate=False)
    assert np.array_equal(exp.masks, le.lime_masks(60, 4, 0))
    assert np.array_equal(exp.probabilities, full.probabilities)            # de-duplicating the <= 16 distinct masks changes no bit
    assert exp.local_exp == full.local_exp
    uniq = np.unique(exp.masks, axis=0)
    want = {tuple(m): p for m, p in zip(uniq.tolist(), loops.stem_mask_probs(stems, uniq, oracle_predictor, SR))}
    ref_probs = np.array([want[tuple(m)] for m in exp.masks.tolist()])
    assert np.abs(exp.probabilities - ref_probs).max() < TOL
    ref = le.fit_lime(exp.masks, ref_probs)
    assert exp.top_label == ref.top_label
    assert np.abs(np.array([exp.by_feature[n] for n in le.COMPONENT_NAMES_4STEMS])
                  - np.array([ref.by_feature[n] for n in le.COMPONENT_NAMES_4STEMS])).max() < TOL
    assert le.predict_fn_unified(stems.sum(0), predictor).shape == (1, 2)


@pytest.mark.parametrize("normalize,seconds", [(False, 6.0), (True, 6.0), (False, 6.4), (True, 6.4)])
def test_fbp_batch_of_tracks_equals_track_by_track(predictor, normalize, seconds):
    # 6.4 s = 200 hops: the iSTFT output is as long as the track and the baselines ride in the band copies' forward;
    # 6.0 s is ragged (187.5 hops): the baselines take a forward of their own
    # 5 tracks x 13 bands with copies_per_chunk = 8 ... the fixture's chunk holds no full band bank: use a wider engine
    p = B200Predictor.random_init(seed=0, copies_per_chunk=32, max_samples=SR * 8)
    try:
        fbp = FrequencyBandPerturbation(p, sr=SR, preset="high_resolution", attenuation=0.25, transition_mode="rel", transition_rel=0.2,
                                        transition_min_hz=5.0, transition_max_hz=500.0, normalize_loudness=normalize)
        sigs = [synth.synth_track(fam, 2, SR, seconds) for fam in ("REAL", "SUNO", "SUNO_PRO", "UDIO", "ElevenLabs")]
        batch = fbp.compute_importance_batch(sigs)            # 2 tracks per group (2 x 13 <= 32): groups of 2, 2, 1
        assert len(batch) == 5
        for sig, b in zip(sigs, batch):
            one = fbp._compute_component_importance(sig, "mixture")
            assert b.baseline_pred == one.baseline_pred       # same kernels on the same data: not a single bit differs
            assert [x["importance"] for x in b.batch_importances] == [x["importance"] for x in one.batch_importances]
            assert np.array_equal(b.importance_map, one.importance_map)
        assert len({b.baseline_pred for b in batch}) == 5     # the tracks really are different
        n_freq, n_time = p.engine.track_shape()               # the last track of the batch is the engine's current track
        assert (n_freq, n_time) == (1025, 1 + len(sigs[-1]) // 512)
        assert np.abs(p.engine.spectrogram() - one.S).max() == 0
    finally:
        p.close()
