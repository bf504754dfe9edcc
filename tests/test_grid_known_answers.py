"""Known-answer tests for the host-side integer logic."""
import numpy as np
import pytest

from audio_deepfake_explainability_b200 import grid


@pytest.mark.parametrize("sr,L,t_stride,pf,sf,n_expected,pfreq,sfreq", [
    (16000, 1920000, 1024, 5.0, 5.0, 60, 51, 51),
    (16000, 1920000, 512, 5.0, 2.5, 228, 51, 26),
    (16000, 1920000, 256, 5.0, 1.25, 825, 51, 13),
    (44100, 5292000, 1024, 5.0, 5.0, 200, 51, 51),
    (44100, 5292000, 512, 5.0, 2.5, 722, 51, 26),
    (44100, 5292000, 256, 5.0, 1.25, 2775, 51, 13),
    (16000, 1920000, 1024, 20.0, 10.0, 27, 205, 102),    # round(102.5) == 102 (half-to-even)
    (44100, 5292000, 1024, 20.0, 10.0, 90, 205, 102),
])
def test_patch_grid_counts(sr, L, t_stride, pf, sf, n_expected, pfreq, sfreq):
    n_freq, n_time = grid.stft_shape(L, 2048, 512)
    assert n_freq == 1025 and n_time == 1 + L // 512
    assert grid.occlusion_patch_sizes(n_freq, pf, sf) == (pfreq, sfreq)
    w = grid.occlusion_windows(n_freq, n_time, 1024, t_stride, pf, sf)
    assert w.shape == (n_expected, 4) and w.dtype == np.int32
    assert (w[:, 1] <= n_time).all() and (w[:, 3] <= n_freq).all()
    # t-major, f-minor ordering
    key = w[:, 0].astype(np.int64) * 100000 + w[:, 2]
    assert (np.diff(key) > 0).all()


def test_patch_grid_degenerate():
    w = grid.occlusion_windows(17, 5, 1024, 1024, 200.0, 50.0)    # patch larger than the spectrogram
    assert w.tolist() == [[0, 5, 0, 17]]
    w = grid.occlusion_windows(9, 10, 4, 4, 0.0, 0.0)             # percent 0 -> max(1, 0) = 1 bin
    assert (w[:, 3] - w[:, 2] == 1).all() and len(w) == 2 * 9


def test_band_bins_16k_and_44k():
    hr = grid.FREQUENCY_BAND_PRESETS["high_resolution"]
    r16 = grid.band_bin_ranges(hr, 16000, 2048)
    exp16 = [(3, 7), (8, 12), (13, 32), (32, 64), (64, 128), (128, 256), (256, 512), (512, 768), (768, 1024),
             (1024, 1024)]
    for (a, b), (ea, eb) in zip(r16[:10], exp16):
        assert (a, b - 1) == (ea, eb)
    assert r16[10:].tolist() == [[0, 0]] * 3                       # bands above Nyquist are empty
    assert r16[2][1] - 1 == r16[3][0] == 32                        # 250 Hz bin belongs to two bands
    r44 = grid.band_bin_ranges(hr, 44100, 2048)
    exp44 = [(1, 2), (3, 4), (5, 11), (12, 23), (24, 46), (47, 92), (93, 185), (186, 278), (279, 371), (372, 464),
             (465, 557), (558, 743), (744, 975)]
    assert [(int(a), int(b) - 1) for a, b in r44] == exp44


def test_keep_mask_known_values():
    f = np.array([0., 50., 75., 100., 150., 200., 225., 250., 300.])
    m = grid.smooth_band_keep_mask(f, 100.0, 200.0, trans=50.0)
    np.testing.assert_allclose(m, [1, 1, 0.5, 0, 0, 0, 0.5, 1, 1], atol=1e-15)
    assert grid.smooth_band_keep_mask(f, 100.0, 200.0, trans=0.0).tolist() == [1, 1, 1, 0, 0, 0, 1, 1, 1]
    tw = [grid.band_transition_width(lo, hi, "rel", 0.2, 5.0, 500.0, 200.0)
          for lo, hi in grid.FREQUENCY_BAND_PRESETS["high_resolution"]]
    assert tw == [8, 8, 30, 50, 100, 200, 400, 400, 400, 400, 400, 500, 500]
    assert grid.band_transition_width(20, 100, "abs", 0.2, 5.0, 500.0, 200.0) == 200.0


def test_coverage_count_half_stride():
    w = grid.occlusion_windows(1025, 3751, 1024, 512, 5.0, 2.5)
    cnt = np.zeros((1025, 3751), np.int32)
    for t0, t1, f0, f1 in w:
        cnt[f0:f1, t0:t1] += 1
    assert cnt[100, 1000] == 4 and cnt[0, 0] == 1 and cnt[30, 0] == 2 and cnt[0, 600] == 2
    assert cnt[:, 3584:].max() == 0 and cnt[1013:, :].max() == 0   # uncovered tails


def test_shard_range():
    for n in (0, 1, 7, 228, 825):
        for ws in (1, 2, 3, 8):
            parts = [grid.shard_range(n, r, ws) for r in range(ws)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(ws - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def test_band_map_view_equals_the_reference_accumulation():
    """grid.band_map_view (read-only broadcast of the map's one distinct column) against the reference's
    map[(freqs >= low) & (freqs <= high), :] += delta (src/dsp_band_ops.py:652-653), overlapping bands included."""
    import numpy as np
    from audio_deepfake_explainability_b200 import grid

    for preset, sr in (("high_resolution", 16000), ("high_resolution", 44100), ("default", 16000)):
        if preset not in grid.FREQUENCY_BAND_PRESETS:
            continue
        bands = list(grid.FREQUENCY_BAND_PRESETS[preset]) + [(100.0, 900.0)]        # one overlapping band
        rows = grid.band_bin_ranges(bands, sr, 2048)
        d = np.random.default_rng(len(bands) + sr).normal(size=len(bands))
        ref = np.zeros((1025, 37))
        freqs = grid.fft_frequencies(sr, 2048)
        for (lo, hi), dd in zip(bands, d):
            ref[(freqs >= lo) & (freqs <= hi), :] += dd
        view = grid.band_map_view(rows, d, 1025, 37)
        assert view.shape == ref.shape and view.dtype == ref.dtype and np.array_equal(view, ref)
        assert not view.flags.writeable
        assert np.array_equal(np.array(view), ref) and np.array(view).flags.writeable
