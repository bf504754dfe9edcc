"""Parity at BASELINE.json's full sizes (120 s / 16 kHz, n_time 3751): the 228-window half-stride sweep of configs[1], the
825-window quarter-stride grid of configs[3] and the 13-band high_resolution FBP of configs[2], checked through
size-independent properties (round trip, linearity, idempotence, chunking invariance, coverage counts) plus a handful of
windows against the CPU oracle (which needs ~0.3 s per evaluation at this size)."""
import numpy as np
import pytest

from audio_deepfake_explainability_b200 import grid, synth
from audio_deepfake_explainability_b200.engine import Engine
from audio_deepfake_explainability_b200.weights import ALPHA_120S, random_state_dict
from oracle import dsp, loops, spectttra                 # checker only

pytestmark = pytest.mark.gpu
CFG = ALPHA_120S
SR, DUR = 16000, 120.0
TOL = 1e-3


@pytest.fixture(scope="module")
def sd():
    return random_state_dict(CFG, 0)


@pytest.fixture(scope="module")
def track():
    return synth.synth_track("SUNO", 0, SR, DUR)


@pytest.fixture(scope="module")
def eng(sd, track):
    e = Engine(CFG, sd, copies_per_chunk=228, max_samples=len(track))
    e.set_track(track)
    yield e
    e.close()


def test_stft_istft_round_trip_full_length(eng, track):
    n_freq, n_time = eng.track_shape()
    assert (n_freq, n_time) == (1025, 3751) == grid.stft_shape(len(track), 2048, 512)
    # occluding an EMPTY rectangle is the identity: iSTFT(STFT(y)) == y up to fp32 rounding (SURVEY 8c vii)
    y_rt = eng.occluded_audio(np.array([[100, 100, 0, 0]], np.int32))[0]
    assert y_rt.shape == (512 * 3750,) and len(track) == 512 * 3750
    assert np.abs(y_rt - track).max() < 1e-5


def test_half_stride_sweep_properties(eng, sd, track):
    n_freq, n_time = eng.track_shape()
    wins = grid.occlusion_windows(n_freq, n_time, 1024, 512, 5.0, 2.5)
    assert len(wins) == 228 and wins[0].tolist() == [0, 1024, 0, 51] and wins[1].tolist() == [0, 1024, 26, 77]
    base = float(eng.predict(track))
    prob = eng.occlusion_sweep(wins, 0.0)
    assert prob.shape == (228,) and np.isfinite(prob).all()
    # determinism, and chunking invariance at full size: any sub-list gives the same bits as the whole sweep
    assert np.array_equal(prob, eng.occlusion_sweep(wins, 0.0))
    sub = np.array([3, 77, 150, 227])
    assert np.array_equal(prob[sub], eng.occlusion_sweep(wins[sub], 0.0))
    # the baseline evaluated inside the sweep's own device pass is predict_track's, bit for bit, and leaves the windows' bits alone
    prob_b, base_b = eng.occlusion_sweep(wins[:227], 0.0, with_baseline=True)       # 227 + 1 copies: one chunk of 228
    assert np.float32(base_b) == np.float32(eng.predict_track()) == np.float32(base) and np.array_equal(prob_b, prob[:227])
    prob_c, base_c = eng.occlusion_sweep(wins, 0.0, with_baseline=True)             # 228 windows fill the chunk: baseline on its own
    assert np.float32(base_c) == np.float32(base) and np.array_equal(prob_c, prob)
    # idempotence of the mask: an empty rectangle and a rectangle that restores the original value change nothing
    same = eng.occlusion_sweep(np.array([[512, 512, 100, 151]], np.int32), 0.0)
    assert abs(float(same[0]) - np.float32(base)) < 2e-5        # predict(y) vs predict(iSTFT(STFT(y))), like the reference
    # a handful of windows against the oracle at full size
    pred = spectttra.OraclePredictor(sd, CFG, "fp32")
    S = dsp.stft(track).numpy()
    base_ref = pred.predict(track, SR)
    assert abs(base - base_ref) < TOL
    delta = np.float64(np.float32(base)) - prob.astype(np.float64)
    for i in sub:
        t0, t1, f0, f1 = wins[i]
        S_occ = S.copy()
        S_occ[f0:f1, t0:t1] = 0.0
        d_ref = base_ref - pred.predict(dsp.istft(S_occ).numpy(), SR)
        assert abs(delta[i] - d_ref) < TOL, f"window {i}: {delta[i]} vs {d_ref}"
    # reductions: bit-exact with the reference accumulation; coverage count 4 in the interior at half stride (SURVEY 8c iv)
    sal = eng.saliency_map(wins, delta)
    assert np.array_equal(sal, loops.saliency_from_windows(wins, delta, n_freq, n_time))
    cnt = eng.saliency_map(wins, np.ones(len(wins)))        # sum of ones / (count + 1e-8): 1 where covered, 0 elsewhere
    f_cov, t_cov = int(wins[:, 3].max()), int(wins[:, 1].max())
    assert (f_cov, t_cov) == (1013, 3584)                   # 38 x 6 positions: bins >= 1013 and frames >= 3584 are never covered (:707)
    assert np.all(cnt[:f_cov, :t_cov] > 0.999) and np.all(cnt[:, t_cov:] == 0) and np.all(cnt[f_cov:, :] == 0)
    for mode, key, desc in ((0, abs, True), (1, abs, False), (2, float, True), (3, float, False)):
        assert eng.rank(delta, mode).tolist() == sorted(range(len(delta)), key=lambda i: key(delta[i]), reverse=desc)


def test_occlusion_is_linear_in_the_removed_patch(eng, track):
    # iSTFT linearity (SURVEY 8c vi): y - y_occ(w) is the patch-only audio, supported on [t0*hop - 1024, (t1-1)*hop + 1024)
    w = np.array([[1024, 2048, 205, 256], [2048, 3072, 410, 461]], np.int32)
    y_occ = eng.occluded_audio(w)
    y_id = eng.occluded_audio(np.array([[0, 0, 0, 0]], np.int32))[0]
    both = y_id - (y_id - y_occ[0]) - (y_id - y_occ[1])
    # zeroing both rectangles at once == subtracting both patch signals (the rectangles are disjoint)
    S = eng.spectrogram()
    S2 = S.copy()
    for t0, t1, f0, f1 in w:
        S2[f0:f1, t0:t1] = 0
    ref = dsp.istft(S2).numpy()
    assert np.abs(both - ref).max() < 2e-6
    for k, (t0, t1, f0, f1) in enumerate(w):
        diff = y_id - y_occ[k]
        lo, hi = t0 * 512 - 1024, (t1 - 1) * 512 + 1024
        assert np.all(diff[: max(lo, 0)] == 0) and np.all(diff[hi:] == 0)      # untouched samples are bit-identical
        assert np.abs(diff[lo:hi]).max() > 1e-4
    # the top-window reconstruction is that same patch signal cut to the window's span (:472-483)
    aud = eng.window_audio(w[:1])[0]
    assert aud.shape == (1024 * 512,)
    assert np.abs(aud - (y_id - y_occ[0])[1024 * 512: 2048 * 512]).max() < 2e-6


def test_quarter_stride_grid_and_sweep(eng, track):
    n_freq, n_time = eng.track_shape()
    wins = grid.occlusion_windows(n_freq, n_time, 1024, 256, 5.0, 1.25)
    assert len(wins) == 825                                                    # configs[3]: 11 x 75
    prob = eng.occlusion_sweep(wins, 0.0)                                       # 4 chunks of <= 228 copies
    assert np.isfinite(prob).all()
    # the half-stride windows are a subset of the quarter-stride grid: same windows, same bits, whatever the batch
    half = grid.occlusion_windows(n_freq, n_time, 1024, 512, 5.0, 2.5)
    index = {tuple(w): i for i, w in enumerate(wins.tolist())}
    common = [(index[tuple(w)], j) for j, w in enumerate(half.tolist()) if tuple(w) in index]
    assert len(common) >= 100
    p_half = eng.occlusion_sweep(half, 0.0)
    assert all(prob[i] == p_half[j] for i, j in common)


def test_fbp_high_resolution_full_length(eng, sd, track):
    bands = grid.FREQUENCY_BAND_PRESETS["high_resolution"] if hasattr(grid, "FREQUENCY_BAND_PRESETS") else None
    if bands is None:
        from audio_deepfake_explainability_b200.dsp_band_ops import FREQUENCY_BAND_PRESETS
        bands = FREQUENCY_BAND_PRESETS["high_resolution"]
    gains = grid.band_gain_table(bands, SR, 2048, 0.25, "rel", 0.2, 5.0, 500.0, 0.0)
    assert gains.shape == (13, 1025)
    assert np.all(gains[10:] == 1.0)                                            # bands above Nyquist: keep == 1 everywhere
    base = float(eng.predict(track))
    for normalize in (False, True):
        prob = eng.fbp_sweep(gains.astype(np.float32), normalize)
        assert np.isfinite(prob).all()
        assert np.all(np.abs(prob[10:] - np.float32(base)) < 2e-5)              # empty bands change nothing (delta ~ 0)
        assert np.array_equal(prob, eng.fbp_sweep(gains.astype(np.float32), normalize))
    # unit gains reproduce the track: band audio of an all-ones gain row == iSTFT(STFT(y))
    y_one = eng.band_audio(np.ones((1, 1025), np.float32))[0]
    assert np.abs(y_one - track).max() < 1e-5
    # one band against the oracle at full size (attenuating 1-2 kHz)
    pred = spectttra.OraclePredictor(sd, CFG, "fp32")
    ref = loops.fbp_component(track, pred, SR, bands=[bands[5]], attenuation=0.25, transition_mode="rel", transition_rel=0.2,
                              transition_min_hz=5.0, transition_max_hz=500.0, normalize_loudness=False)
    got = np.float64(np.float32(base)) - float(eng.fbp_sweep(gains[5:6].astype(np.float32), False)[0])
    assert abs(got - ref.batch_importances[0]["importance"]) < TOL
