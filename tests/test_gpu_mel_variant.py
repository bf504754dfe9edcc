"""Mel-domain explainer variant (spec_type: mel, src/spectrogram_explainability.py:367-377, 394-402) against the
BUILDER-DEFINED oracle (oracle/mel.py: Slaney melspectrogram restated exactly; clipped-least-squares + projected-gradient
NNLS and hashed-phase Griffin-Lim in place of librosa's L-BFGS-B / unseeded RNG - no reference parity can exist for those)."""
import numpy as np
import pytest

from audio_deepfake_explainability_b200 import grid, mel_host, synth
from audio_deepfake_explainability_b200.dsp_band_ops import FrequencyBandPerturbation
from audio_deepfake_explainability_b200.sonics_api import B200Predictor
from audio_deepfake_explainability_b200.spectrogram_explainability import SpectrogramExplainability
from audio_deepfake_explainability_b200.weights import ALPHA_120S, random_state_dict
from oracle import mel as omel, spectttra                  # checker only

pytestmark = pytest.mark.gpu
SR, N_MELS = 16000, 128


@pytest.fixture(scope="module")
def predictor():
    p = B200Predictor.random_init(0, copies_per_chunk=8, max_samples=16000 * 12)
    yield p
    p.close()


@pytest.fixture(scope="module")
def track():
    return synth.synth_track("UDIO", 1, SR, 10.0)


@pytest.fixture(scope="module")
def basis(predictor):
    A = mel_host.mel_filterbank(SR, 2048, N_MELS)
    P, step = mel_host.nnls_operators(A)
    predictor.engine.set_mel_basis(A, P, step)
    return A


def test_power_mel_spectrogram_matches_librosa_restatement(predictor, track, basis):
    eng = predictor.engine
    eng.set_track(track)
    M = eng.mel_spectrogram()
    ref = omel.melspectrogram(track, SR, 2048, 512, 2048, N_MELS)
    assert M.shape == ref.shape == (N_MELS, 1 + len(track) // 512)
    assert np.abs(M - ref).max() <= 2e-5 * ref.max()


@pytest.mark.parametrize("nnls_iter", [0, 16])
def test_nnls_and_griffin_lim_match_the_builder_oracle(predictor, track, basis, nnls_iter):
    eng = predictor.engine
    eng.set_track(track)
    M = omel.melspectrogram(track, SR, 2048, 512, 2048, N_MELS)
    # n_iter = 0: audio = istft(mag * initial phases): isolates NNLS + phase hash
    y0 = eng.mel_sweep(eng.MASK_NONE, None, 0, nnls_iter, seed=5, first_index=2, want_prob=False, want_audio=True)[0]
    r0 = omel.mel_to_audio_builder(M, SR, n_iter=0, seed=5, index=2, nnls_iter=nnls_iter)
    assert y0.shape == r0.shape
    scale = np.abs(r0).max()
    assert np.abs(y0 - r0).max() <= 2e-4 * scale
    # a few Griffin-Lim iterations (fast GL, momentum 0.99)
    y4 = eng.mel_sweep(eng.MASK_NONE, None, 4, nnls_iter, seed=5, first_index=2, want_prob=False, want_audio=True)[0]
    r4 = omel.mel_to_audio_builder(M, SR, n_iter=4, seed=5, index=2, nnls_iter=nnls_iter)
    assert np.abs(y4 - r4).max() <= 2e-3 * np.abs(r4).max()
    assert np.abs(y4 - y0).max() > 1e-2 * scale              # the iterations do something
    # different phase index -> different audio; same index -> same bits
    assert not np.array_equal(y0, eng.mel_sweep(eng.MASK_NONE, None, 0, nnls_iter, seed=5, first_index=3, want_prob=False, want_audio=True)[0])
    assert np.array_equal(y0, eng.mel_sweep(eng.MASK_NONE, None, 0, nnls_iter, seed=5, first_index=2, want_prob=False, want_audio=True)[0])


def test_mel_occlusion_map_matches_the_oracle_loop(predictor, track):
    sd = random_state_dict(ALPHA_120S, 0)
    opred = spectttra.OraclePredictor(sd, ALPHA_120S, "fp32")
    kw = dict(patch_time_frames=128, stride_time_frames=128, patch_freq_percent=50.0, stride_freq_percent=50.0)
    ex = SpectrogramExplainability(predictor, sr=SR, duration=120, n_mels=N_MELS, n_iter=4, spec_type="mel", method="occlusion",
                                   use_original_audio=False, top_n_windows=2, mel_seed=11, **kw)
    res = ex.occlusion_map_from_wave(track, baseline_threshold=0.0, verbose=False)
    ref = omel.occlusion_map_mel(track, opred, SR, n_mels=N_MELS, n_iter=4, baseline_threshold=0.0, seed=11, **kw)
    assert res.S.shape == ref.S.shape == (N_MELS, 313)
    assert len(res.patch_importances) == len(ref.patch_importances) == 4
    for a, b in zip(res.patch_importances, ref.patch_importances):
        assert (a["t_start"], a["t_end"], a["f_start"], a["f_end"]) == (b["t_start"], b["t_end"], b["f_start"], b["f_end"])
        assert abs(a["importance"] - b["importance"]) < 2e-3, (a, b)
    assert abs(res.baseline_pred - ref.baseline_pred) < 1e-3
    assert res.importance_map.shape == (N_MELS, 313)
    imp = np.array([p["importance"] for p in res.patch_importances])
    from oracle import loops
    assert np.array_equal(res.importance_map, loops.saliency_from_windows(
        np.array([[p["t_start"], p["t_end"], p["f_start"], p["f_end"]] for p in res.patch_importances]), imp, N_MELS, 313))
    # top-window audio: keep-only patch inverted and sliced to the window's span (:472-483)
    groups = ex.top_window_groups(res.patch_importances, 2, "t")
    segs = ex.window_audio(track, groups["best"]["windows"])
    assert len(segs) == 2 and all(len(s) == 128 * 512 for s in segs) and all(np.isfinite(s).all() for s in segs)


def test_fbp_mel_is_builder_defined_but_consistent(predictor, track):
    sd = random_state_dict(ALPHA_120S, 0)
    opred = spectttra.OraclePredictor(sd, ALPHA_120S, "fp32")
    bands = [(100, 250), (1000, 2000), (4000, 6000)]
    fbp = FrequencyBandPerturbation(predictor, preset="t", presets={"t": bands}, attenuation=0.25, transition_mode="rel", transition_rel=0.2,
                                    transition_min_hz=5.0, transition_max_hz=500.0, sr=SR, n_mels=N_MELS, n_iter=4, spec_type="mel",
                                    normalize_loudness=False, mel_seed=3)
    r = fbp._compute_component_importance(track, "mixture")
    assert r.importance_map.shape == (N_MELS, 313) and len(r.batch_importances) == 3
    M = omel.melspectrogram(track, SR, 2048, 512, 2048, N_MELS)
    gains = omel.mel_band_gains(bands, SR, N_MELS, 0.25, "rel", 0.2, 5.0, 500.0, 0.0)
    base = opred.predict(track, SR)
    for i, b in enumerate(r.batch_importances):
        y_p = omel.mel_to_audio_builder((M * gains[i][:, None]).astype(np.float32), SR, n_iter=4, seed=3, index=i)
        y_p = np.pad(y_p, (0, len(track) - len(y_p)))
        assert abs(b["importance"] - (base - opred.predict(y_p, SR))) < 2e-3
