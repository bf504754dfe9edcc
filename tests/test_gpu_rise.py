"""RISE (src/spectrogram_explainability.py:722-806) on the GPU: the keep masks are generated inside the iSTFT load stage
and re-derived in the map reduction from a counter-based hash; the CPU oracle restates the same hash, so masked audio and
the accumulated map are compared bit-for-bit / to fp32 rounding, predictions within the north_star tolerance."""
import numpy as np
import pytest

from audio_deepfake_explainability_b200 import synth
from audio_deepfake_explainability_b200.sonics_api import B200Predictor
from audio_deepfake_explainability_b200.spectrogram_explainability import RiseResult, SpectrogramExplainability
from audio_deepfake_explainability_b200.weights import ALPHA_120S, random_state_dict
from oracle import dsp, loops, spectttra                      # checker only

pytestmark = pytest.mark.gpu
SR, TOL = 16000, 1e-3


@pytest.fixture(scope="module")
def predictor():
    p = B200Predictor.random_init(seed=0, copies_per_chunk=4, max_samples=SR * 12)
    yield p
    p.close()


@pytest.fixture(scope="module")
def track():
    return synth.synth_track("ElevenLabs", 1, SR, 8.0)


@pytest.mark.parametrize("p_keep", [0.5, 0.9])
def test_masked_audio_uses_the_oracle_mask_bits(predictor, track, p_keep):
    eng = predictor.engine
    eng.set_track(track)
    n_freq, n_time = eng.track_shape()
    S = dsp.stft(track).numpy()
    audio = eng.rise_audio(3, seed=11, keep_probability=p_keep, first_mask=5)
    for i in range(3):
        mask = loops.rise_keep_mask(11, 5 + i, n_freq, n_time, p_keep)
        assert abs(mask.mean() - p_keep) < 0.01
        ref = dsp.istft(S * mask.astype(np.float32)).numpy()
        assert audio[i].shape == ref.shape
        assert np.abs(audio[i] - ref).max() < 5e-6 * max(1.0, np.abs(ref).max())      # a single wrong bit would show at ~1e-3


def test_rise_map_is_bit_exact_with_the_reference_accumulation(predictor, track):
    eng = predictor.engine
    eng.set_track(track)
    n_freq, n_time = eng.track_shape()
    rng = np.random.default_rng(0)
    preds = rng.uniform(0.2, 0.9, 37)
    got = eng.rise_map(preds, seed=3, keep_probability=0.5)
    want = np.zeros((n_freq, n_time))
    for i, p in enumerate(preds):
        want += loops.rise_keep_mask(3, i, n_freq, n_time, 0.5).astype(float) * float(p)      # :783
    want = want / (len(preds) * 0.5 + 1e-8)                                                   # :798
    assert np.array_equal(got, want)
    assert np.array_equal(eng.rise_map(np.zeros(0), 3, 0.5), np.zeros((n_freq, n_time)))       # no masks: 0 / 1e-8


def test_rise_explainer_matches_oracle_loop(predictor, track):
    ex = SpectrogramExplainability(predictor, sr=SR, spec_type="stft", method="rise", n_masks=10, mask_probability=0.5, rise_seed=7)
    res = ex.rise_map_from_wave(track, baseline_threshold=0.0, verbose=False)
    assert isinstance(res, RiseResult) and res.importance_map.shape == res.S.shape
    oracle = spectttra.OraclePredictor(random_state_dict(ALPHA_120S, 0), ALPHA_120S, "fp32")
    ref = loops.rise_map(track, oracle, SR, n_masks=10, mask_probability=0.5, seed=7, baseline_threshold=0.0)
    assert abs(res.baseline_pred - ref.baseline_pred) < TOL
    probs = predictor.engine.rise_sweep(10, 7, 0.5)
    assert np.abs(probs - np.array(ref.predictions)).max() < TOL
    assert np.ptp(ref.predictions) > 1e-4                                   # not vacuous
    assert res.importance_map.min() == 0.0 and abs(res.importance_map.max() - 1.0) < 1e-6
    # the min-max scaling divides by the map's range, which amplifies the 1e-3 tolerance on the predictions accordingly
    raw = predictor.engine.rise_map(np.array(ref.predictions), 7, 0.5)
    assert np.array_equal(raw, ref.raw_map)                                  # same predictions -> bit-identical map
    # sharding invariance: masks 4..9 alone give the same bits as inside the full sweep
    assert np.array_equal(probs[4:], predictor.engine.rise_sweep(6, 7, 0.5, first_mask=4))
    # skipped below the threshold (:741-744), and the occlusion entry point refuses a RISE explainer loudly
    assert ex.rise_map_from_wave(track, baseline_threshold=1.1, verbose=False).importance_map is None
    with pytest.raises(NotImplementedError):
        ex.occlusion_map_from_wave(track)
