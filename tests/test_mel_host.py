"""Mel-domain variant, host side: the Slaney filterbank against torchaudio's independent implementation of the same
definition (melscale_fbanks(norm='slaney', mel_scale='slaney')) and against the oracle's restatement of librosa.filters.mel;
the builder's NNLS; the FBP-mel band gains."""
import numpy as np
import pytest

from audio_deepfake_explainability_b200 import grid, mel_host
from oracle import mel as omel


@pytest.mark.parametrize("sr,n_mels", [(16000, 128), (44100, 128), (44100, 512), (22050, 64)])
def test_slaney_filterbank_matches_torchaudio_and_the_oracle(sr, n_mels):
    import torchaudio
    A = mel_host.mel_filterbank(sr, 2048, n_mels)
    assert A.shape == (n_mels, 1025) and A.dtype == np.float32
    assert np.array_equal(A, omel.slaney_mel_filterbank(sr, 2048, n_mels))
    ref = torchaudio.functional.melscale_fbanks(1025, 0.0, sr / 2, n_mels, sr, norm="slaney", mel_scale="slaney").numpy().T
    assert np.abs(A - ref).max() <= 2e-6 * max(1.0, np.abs(ref).max())
    # every FFT bin feeds at most two, adjacent, filters (what the CUDA kernels rely on)
    nz = A != 0
    assert nz.sum(0).max() <= 2
    for k in np.nonzero(nz.sum(0) == 2)[0]:
        i = np.nonzero(nz[:, k])[0]
        assert i[1] == i[0] + 1


def test_mel_scale_round_trip_and_known_points():
    f = np.array([0.0, 200.0, 1000.0, 4000.0, 8000.0])
    assert np.allclose(mel_host.mel_to_hz(mel_host.hz_to_mel(f)), f)
    assert np.isclose(mel_host.hz_to_mel(1000.0), 15.0) and np.isclose(mel_host.hz_to_mel(200.0), 3.0)


def test_builder_nnls_is_nonnegative_and_reduces_the_residual():
    rng = np.random.default_rng(0)
    A = mel_host.mel_filterbank(16000, 2048, 128)
    X_true = rng.random((1025, 6)).astype(np.float32) ** 4
    B = A @ X_true
    P, step = mel_host.nnls_operators(A)
    P2, step2 = omel.nnls_operators(A)
    assert np.array_equal(P, P2) and np.isclose(step, step2)
    X0 = np.maximum(P @ B, 0)
    X = omel.nnls_builder(A, B, 16)
    assert X.min() >= 0
    assert np.linalg.norm(A @ X - B) <= np.linalg.norm(A @ X0 - B) + 1e-6
    assert np.linalg.norm(A @ X - B) < 0.05 * np.linalg.norm(B)


def test_fbp_mel_gains_match_the_oracle_and_cover_the_bands():
    bands = grid.FREQUENCY_BAND_PRESETS["high_resolution"]
    g = mel_host.mel_band_gain_table(bands, 16000, 128, 0.25, "rel", 0.2, 5.0, 500.0, 0.0)
    assert np.array_equal(g, omel.mel_band_gains(bands, 16000, 128, 0.25, "rel", 0.2, 5.0, 500.0, 0.0))
    assert g.shape == (13, 128) and np.isclose(g.min(), 0.25) and g.max() == 1.0
    assert np.all(g[10:] == 1.0)                               # bands above Nyquist change nothing
    rows = mel_host.mel_band_rows(bands, 16000, 128)
    assert len(rows[0]) >= 1 and len(rows[12]) == 0
    c = mel_host.mel_band_centres(16000, 128)
    assert np.all(np.diff(c) > 0) and c[0] > 0 and c[-1] < 8000


def test_phase_hash_is_the_rise_hash():
    from oracle import loops
    u = omel.phase_uniform(7, 3, 1025, 5)
    assert u.shape == (1025, 5) and u.min() >= 0 and u.max() <= 1
    # the same (seed, index, cell) -> bit stream that the RISE keep masks threshold
    keep = loops.rise_keep_mask(7, 3, 1025, 5, 0.5)
    assert np.array_equal(keep, u.astype(np.float64) * 4294967296.0 < 2147483648.0) or np.mean(keep == (u < 0.5)) > 0.9999
