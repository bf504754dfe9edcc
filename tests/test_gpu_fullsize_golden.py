"""Full-size parity against CACHED oracle sweeps (tests/golden/fullsize_*.npz, shipped_config_44k_*.npz; made by
oracle/make_golden_fullsize.py in the build container - the CPU oracle needs ~0.5 s per evaluation at these sizes):

* every one of the 228 windows of the configs[1] sweep, on the bench track (REAL 0) and on a band-limited track (SUNO 0),
  against the fp32 oracle (the reference's arithmetic) and the bf16-GEMM-input oracle (the engine's arithmetic contract);
* the four top-k groups (src/spectrogram_explainability.py:428-434, 566-587) derived from the engine's deltas against the
  groups derived from the oracle's: the SET of every group must be identical whenever the oracle's margin at the k-th
  boundary exceeds twice the measured per-window error; members may differ only among windows whose keys lie within that
  error of the boundary; with ``tie_epsilon`` (grid.snap_ties) the groups are identical outright;
* all 13 bands of the high_resolution FBP bank, both normalize_loudness settings (configs[2]);
* the reference's own shipped configuration: 44.1 kHz, n_time 10 336, 1024 x 20 % windows at 10 % stride -> 90 windows.
"""
from pathlib import Path

import numpy as np
import pytest

from audio_deepfake_explainability_b200 import grid, synth
from audio_deepfake_explainability_b200.engine import Engine
from audio_deepfake_explainability_b200.weights import ALPHA_120S, random_state_dict
from oracle import loops                                  # checker only

pytestmark = pytest.mark.gpu
CFG = ALPHA_120S
GOLDEN = Path(__file__).parent / "golden"
TOL = 1e-3                                                # north_star: per-window delta-prob within 1e-3 absolute
TOP_N = 5
TIE_EPS = 1e-4


@pytest.fixture(scope="module")
def sd():
    return random_state_dict(CFG, 0)


def _group_keys(delta, group):
    return np.abs(delta) if group in ("best", "worst") else delta


def _check_groups(d_eng, d_ref, err, label):
    """Set equality under the margin rule; returns a printable summary."""
    g_eng, g_ref = grid.topk_window_groups(d_eng, TOP_N), grid.topk_window_groups(d_ref, TOP_N)
    margins = grid.topk_boundary_margins(d_ref, TOP_N)
    summary = {}
    for g in ("best", "worst", "most_influential"):
        a, b = set(g_eng[g].tolist()), set(g_ref[g].tolist())
        if margins[g] > 2 * err:
            assert a == b, f"{label} {g}: engine {sorted(a)} vs oracle {sorted(b)} with margin {margins[g]:.2e} > 2 x err {err:.2e}"
        else:
            # inside the noise band the sets may trade members, but only windows whose key is within 2 err of a traded one
            keys = np.abs(d_ref)
            for i in a ^ b:
                others = (b - a) if i in a else (a - b)
                assert any(abs(keys[i] - keys[j]) <= 2 * err + 1e-12 for j in others), f"{label} {g}: window {i} traded outside the noise band"
        # rank order: positions may differ only between windows whose oracle keys are closer than 2 err
        if a == b:
            for x, y in zip(g_eng[g], g_ref[g]):
                if x != y:
                    assert abs(abs(d_ref[x]) - abs(d_ref[y])) <= 2 * err + 1e-12, f"{label} {g}: order differs beyond the noise band"
        summary[g] = {"set_equal": a == b, "order_equal": bool(np.array_equal(g_eng[g], g_ref[g])), "margin": margins[g]}
    return summary


@pytest.mark.parametrize("family", ["REAL", "SUNO"])
def test_all_228_windows_and_topk_groups_against_cached_oracle(sd, family):
    z = np.load(GOLDEN / f"fullsize_occlusion_{family}0.npz")
    y = synth.synth_track(family, 0, 16000, 120.0)
    eng = Engine(CFG, sd, copies_per_chunk=229, max_samples=len(y))
    try:
        eng.set_track(y)
        n_freq, n_time = eng.track_shape()
        wins = grid.occlusion_windows(n_freq, n_time, 1024, 512, 5.0, 2.5)
        assert np.array_equal(wins, z["windows"])
        prob, base = eng.occlusion_sweep(wins, 0.0, with_baseline=True)
        d_eng = np.float64(np.float32(base)) - prob.astype(np.float64)
        report = {}
        for mode in ("fp32", "bf16"):
            d_ref = z[f"delta_{mode}"]
            assert abs(float(base) - float(z[f"base_{mode}"])) < TOL
            err = float(np.abs(d_eng - d_ref).max())
            assert err < TOL, f"{family} vs {mode} oracle: max |delta error| {err:.3e}"
            big = np.abs(d_ref) > 1e-3
            rel = float((np.abs(d_eng - d_ref)[big] / np.abs(d_ref[big])).max()) if big.any() else 0.0
            report[mode] = {"max_abs_err": err, "max_rel_err(|d|>1e-3)": rel, "groups": _check_groups(d_eng, d_ref, err, f"{family}/{mode}")}
            # with the tie epsilon both sides agree on the groups when the snapped keys are separated by more than the error
            s_eng, s_ref = grid.snap_ties(d_eng, TIE_EPS), grid.snap_ties(d_ref, TIE_EPS)
            _check_groups(s_eng, s_ref, err, f"{family}/{mode}/tie_eps")
        print(f"\n[fullsize golden] {family}: {report}")
        # saliency of the engine deltas == the reference accumulation of the same deltas, bit for bit (:695-696, :707)
        sal = eng.saliency_map(wins, d_eng)
        assert np.array_equal(sal, loops.saliency_from_windows(wins, d_eng, n_freq, n_time))
        # and the map built from oracle deltas differs by at most the per-window error (a mean of <= 4 windows per cell)
        sal_ref = loops.saliency_from_windows(wins, z["delta_fp32"], n_freq, n_time)
        assert np.abs(sal - sal_ref).max() < TOL
    finally:
        eng.close()


def test_band_limited_track_has_sub_noise_ties_that_tie_epsilon_resolves():
    """SUNO 0 has no energy above 5 kHz: windows over those bins change nothing.  Their |delta| is the arithmetic noise of
    the classifier; the raw 'worst' order among them is arithmetic-specific, the snapped order is grid order everywhere."""
    z = np.load(GOLDEN / "fullsize_occlusion_SUNO0.npz")
    a, b = z["delta_fp32"], z["delta_bf16"]
    quiet = np.abs(a) < TIE_EPS
    assert quiet.sum() >= TOP_N                           # more sub-noise windows than the group holds
    worst = grid.topk_window_groups(a, TOP_N, TIE_EPS)["worst"]
    assert worst.tolist() == np.nonzero(np.abs(grid.snap_ties(a, TIE_EPS)) == 0)[0][:TOP_N].tolist()     # grid order
    if (np.abs(b) < TIE_EPS).sum() >= TOP_N and np.array_equal(np.abs(a) < TIE_EPS, np.abs(b) < TIE_EPS):
        assert np.array_equal(worst, grid.topk_window_groups(b, TOP_N, TIE_EPS)["worst"])


def test_fbp_all_13_bands_both_loudness_settings(sd):
    from audio_deepfake_explainability_b200.dsp_band_ops import FREQUENCY_BAND_PRESETS
    z = np.load(GOLDEN / "fullsize_fbp_SUNO0.npz")
    y = synth.synth_track("SUNO", 0, 16000, 120.0)
    bands = FREQUENCY_BAND_PRESETS["high_resolution"]
    assert np.array_equal(np.asarray(bands), z["bands"])
    gains = grid.band_gain_table(bands, 16000, 2048, 0.25, "rel", 0.2, 5.0, 500.0, 0.0).astype(np.float32)
    eng = Engine(CFG, sd, copies_per_chunk=16, max_samples=len(y))
    try:
        eng.set_track(y)
        base = float(eng.predict_track())
        for normalize in (False, True):
            assert abs(base - float(z[f"base_norm{int(normalize)}"])) < TOL
            prob = eng.fbp_sweep(gains, normalize)
            d_eng = np.float64(np.float32(base)) - prob.astype(np.float64)
            err = np.abs(d_eng - z[f"delta_norm{int(normalize)}"])
            print(f"\n[fullsize golden] FBP normalize={normalize}: max |delta error| {err.max():.3e} over 13 bands")
            assert err.max() < TOL, f"normalize={normalize}: {err}"
    finally:
        eng.close()


def test_reference_shipped_configuration_44k(sd):
    """configs/Spec_occlusion_configs/spectrogram_explainability.yaml:35-63: sr 44 100 (the classifier ignores sr and resizes
    the 10 336 mel frames to 3744, src/sonics_api.py:268-271), 1024 x 20 % windows at 10 % stride."""
    z = np.load(GOLDEN / "shipped_config_44k_UDIO0.npz")
    y = synth.synth_track("UDIO", 0, 44100, 120.0)
    assert len(y) == 5292000
    eng = Engine(CFG, sd, copies_per_chunk=91, max_samples=len(y))
    try:
        eng.set_track(y)
        n_freq, n_time = eng.track_shape()
        assert (n_freq, n_time) == (1025, 10336)
        wins = grid.occlusion_windows(n_freq, n_time, 1024, 1024, 20.0, 10.0)
        assert len(wins) == 90 and np.array_equal(wins, z["windows"])
        prob, base = eng.occlusion_sweep(wins, 0.0, with_baseline=True)
        assert abs(float(base) - float(z["base_fp32"])) < TOL
        d_eng = np.float64(np.float32(base)) - prob.astype(np.float64)
        pick = z["pick"]
        err = np.abs(d_eng[pick] - z["delta_fp32"])
        print(f"\n[fullsize golden] shipped 44.1 kHz config: max |delta error| {err.max():.3e} over {len(pick)} windows")
        assert err.max() < TOL, err
        # iSTFT output is 480 samples shorter than the track (10 335 x 512 = 5 291 520): zero padded (:679-680); the identity
        # occlusion reproduces the track on the covered samples
        y_rt = eng.occluded_audio(np.array([[0, 0, 0, 0]], np.int32))[0]
        assert y_rt.shape == (5291520,) and np.abs(y_rt - y[:5291520]).max() < 1e-5
        # the full map: bit-exact accumulation, coverage of the 10 x 9 grid (frames < 10 240, bins < 1021)
        sal = eng.saliency_map(wins, d_eng)
        assert np.array_equal(sal, loops.saliency_from_windows(wins, d_eng, n_freq, n_time))
        cnt = eng.saliency_map(wins, np.ones(len(wins)))
        assert np.all(cnt[:1021, :10240] > 0.999) and np.all(cnt[:, 10240:] == 0) and np.all(cnt[1021:, :] == 0)
        # sparse path == dense path at this size too: a sub-list gives the same bits
        sub = np.array([0, 44, 89])
        assert np.array_equal(prob[sub], eng.occlusion_sweep(wins[sub], 0.0))
    finally:
        eng.close()
