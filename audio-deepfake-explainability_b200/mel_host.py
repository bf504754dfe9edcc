"""Host-side operators of the mel-domain explainer variant (``spec_type: mel``).

Tiny float64 linear algebra that the reference gets from librosa on the host and that stays on the host here too (it runs
once per (sr, n_fft, n_mels) configuration): the Slaney mel filterbank of ``librosa.feature.melspectrogram`` /
``mel_to_stft`` (``librosa.filters.mel(htk=False, norm='slaney')``; src/spectrogram_explainability.py:367-377, 394-402), its
pseudo-inverse and Lipschitz constant for the builder-defined NNLS that the CUDA kernels iterate (csrc/mel_domain.cu), the
mel-bin window grid and the FBP-mel band gains.  Everything per perturbed copy runs on the GPU.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np

from . import grid

_F_SP = 200.0 / 3
_MIN_LOG_HZ = 1000.0
_MIN_LOG_MEL = _MIN_LOG_HZ / _F_SP
_LOGSTEP = np.log(6.4) / 27.0


def hz_to_mel(freq) -> np.ndarray:
    """Slaney (auditory toolbox) mel scale, ``librosa.hz_to_mel(htk=False)``: linear below 1 kHz, logarithmic above."""
    freq = np.asarray(freq, dtype=np.float64)
    log_part = _MIN_LOG_MEL + np.log(np.maximum(freq, _MIN_LOG_HZ) / _MIN_LOG_HZ) / _LOGSTEP
    return np.where(freq >= _MIN_LOG_HZ, log_part, freq / _F_SP)


def mel_to_hz(mels) -> np.ndarray:
    mels = np.asarray(mels, dtype=np.float64)
    return np.where(mels >= _MIN_LOG_MEL, _MIN_LOG_HZ * np.exp(_LOGSTEP * (mels - _MIN_LOG_MEL)), _F_SP * mels)


def mel_edge_frequencies(n_mels: int, fmin: float, fmax: float) -> np.ndarray:
    """``n_mels + 2`` band-edge frequencies, equally spaced on the mel scale (``librosa.mel_frequencies``)."""
    return mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2))


def mel_filterbank(sr: float, n_fft: int, n_mels: int, fmin: float = 0.0, fmax: Optional[float] = None) -> np.ndarray:
    """float32 ``[n_mels, 1 + n_fft//2]`` triangular filters with Slaney area normalisation (``librosa.filters.mel``)."""
    fmax = float(sr) / 2 if fmax is None else float(fmax)
    edges = mel_edge_frequencies(n_mels, fmin, fmax)
    bins = np.fft.rfftfreq(n_fft, 1.0 / sr)
    rising = (bins[None, :] - edges[:-2, None]) / (edges[1:-1] - edges[:-2])[:, None]
    falling = (edges[2:, None] - bins[None, :]) / (edges[2:] - edges[1:-1])[:, None]
    tri = np.maximum(0.0, np.minimum(rising, falling)).astype(np.float32)
    # librosa scales its float32 weights in place by the float64 area norm: one rounding, of the float64 product
    return (tri.astype(np.float64) * (2.0 / (edges[2:] - edges[:-2]))[:, None]).astype(np.float32)


def nnls_operators(basis: np.ndarray) -> Tuple[np.ndarray, float]:
    """(``pinv(basis)`` float32 ``[n_freq, n_mels]``, ``1 / ||basis||_2^2``): starting point and step of the builder's NNLS."""
    b64 = np.asarray(basis, dtype=np.float64)
    return np.linalg.pinv(b64).astype(np.float32), float(np.float32(1.0 / np.linalg.norm(b64, 2) ** 2))


def power_to_db_refmax(S: np.ndarray) -> np.ndarray:
    """``librosa.power_to_db(S, ref=np.max)`` (visualisation only, :377)."""
    S = np.asarray(S, dtype=np.float32)
    out = 10.0 * np.log10(np.maximum(1e-10, S)) - 10.0 * np.log10(max(1e-10, float(S.max())))
    return np.maximum(out, out.max() - 80.0)


def mel_band_gain_table(bands: Sequence[Tuple[float, float]], sr: float, n_mels: int, attenuation: float, transition_mode: str,
                        transition_rel: float, transition_min_hz: float, transition_max_hz: float, transition_hz: float,
                        fmax: Optional[float] = None) -> np.ndarray:
    """FBP-mel gains float64 ``[n_bands, n_mels]``: the raised-cosine keep mask of src/dsp_band_ops.py:236-259, 576 evaluated
    at every mel bin's CENTRE frequency.  (Builder-defined: the reference rejects spec_type='mel' for FBP, :357-359.)"""
    centres = mel_band_centres(sr, n_mels, fmax)
    out = np.empty((len(bands), n_mels))
    for b, (low, high) in enumerate(bands):
        trans = grid.band_transition_width(low, high, transition_mode, transition_rel, transition_min_hz, transition_max_hz, transition_hz)
        keep = grid.smooth_band_keep_mask(centres, low, high, trans)
        out[b] = keep + attenuation * (1.0 - keep)
    return out


def mel_band_centres(sr: float, n_mels: int, fmax: Optional[float] = None) -> np.ndarray:
    return mel_edge_frequencies(n_mels, 0.0, float(sr) / 2 if fmax is None else float(fmax))[1:-1]


def mel_band_rows(bands: Sequence[Tuple[float, float]], sr: float, n_mels: int, fmax: Optional[float] = None) -> list:
    """Index arrays of the mel bins whose centre lies in ``[low, high]`` (hard, inclusive edges like :652-653)."""
    c = mel_band_centres(sr, n_mels, fmax)
    return [np.nonzero((c >= low) & (c <= high))[0] for (low, high) in bands]
