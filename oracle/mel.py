"""ORACLE (test infrastructure only - never imported by the product path): the mel-domain explainer variant.

Reference: ``spec_type: mel`` of ``SpectrogramExplainability`` (src/spectrogram_explainability.py:367-377 forward,
:394-402 inverse, loop :663-703 unchanged) and the builder-defined FBP-mel (the reference rejects it,
src/dsp_band_ops.py:357-359).  The arithmetic lives in un-vendored, un-pinned ``librosa``:

  forward  librosa.feature.melspectrogram(y, sr, n_mels, n_fft, hop_length, win_length, fmax)
             = filters.mel(sr, n_fft, n_mels, fmin=0, fmax, htk=False, norm='slaney') @ |stft(y)|^2        (restated exactly)
  inverse  librosa.feature.inverse.mel_to_audio(S, sr, n_fft, hop_length, win_length, n_iter)
             = griffinlim(mel_to_stft(S) , n_iter, momentum=0.99, init='random', random_state=None)

PARITY STATUS: **builder-defined, no reference parity can exist.**  The reference's inverse is not reproducible even
reference-vs-reference: Griffin-Lim starts from UNSEEDED random phases, and ``mel_to_stft`` runs scipy's L-BFGS-B NNLS
whose iterates depend on the LAPACK build.  The build therefore defines (and this file restates for the CPU):
  * NNLS:  X0 = max(0, pinv(A) B) - librosa's own starting point (util.nnls) - followed by ``nnls_iter`` projected-gradient
           steps X <- max(0, X - (1/L) A^T (A X - B)), L = ||A||_2^2, instead of L-BFGS-B;
  * phases: angle(cell) = 2 pi u with u the counter-based hash of (seed, copy index, cell) shared with the RISE masks
           (csrc/common.h: rise_mask_key / hash_lowbias32), instead of np.random;
  * the Griffin-Lim recursion itself is librosa's, line for line (fast Griffin-Lim, momentum 0.99, eps = tiny(float32)).
The Slaney filterbank, the power mel spectrogram and the loop / indexing semantics are restated exactly.
"""
from __future__ import annotations

from typing import List, NamedTuple, Optional, Sequence, Tuple

import numpy as np
import torch

from . import dsp
from .loops import _lowbias32

F_SP = 200.0 / 3
MIN_LOG_HZ = 1000.0
MIN_LOG_MEL = MIN_LOG_HZ / F_SP
LOGSTEP = np.log(6.4) / 27.0


def hz_to_mel(f):
    """librosa.hz_to_mel(htk=False): Slaney's auditory-toolbox scale (linear below 1 kHz, log above)."""
    f = np.asarray(f, dtype=np.float64)
    lin = f / F_SP
    return np.where(f >= MIN_LOG_HZ, MIN_LOG_MEL + np.log(np.maximum(f, MIN_LOG_HZ) / MIN_LOG_HZ) / LOGSTEP, lin)


def mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    return np.where(m >= MIN_LOG_MEL, MIN_LOG_HZ * np.exp(LOGSTEP * (m - MIN_LOG_MEL)), F_SP * m)


def mel_frequencies(n_mels: int, fmin: float, fmax: float) -> np.ndarray:
    return mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels))


def slaney_mel_filterbank(sr: float, n_fft: int, n_mels: int, fmin: float = 0.0, fmax: Optional[float] = None) -> np.ndarray:
    """librosa.filters.mel(sr=, n_fft=, n_mels=, fmin=0, fmax=None -> sr/2, htk=False, norm='slaney', dtype=float32)."""
    fmax = float(sr) / 2 if fmax is None else float(fmax)
    fftfreqs = np.fft.rfftfreq(n_fft, 1.0 / sr)
    mel_f = mel_frequencies(n_mels + 2, fmin, fmax)
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    weights = np.zeros((n_mels, 1 + n_fft // 2), dtype=np.float32)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2: n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]                    # in place: float32(float64 product), like librosa
    return weights


def melspectrogram(y, sr: int, n_fft: int = 2048, hop_length: int = 512, win_length: int = 2048, n_mels: int = 128,
                   fmax: Optional[float] = None) -> np.ndarray:
    """float32 ``[n_mels, 1 + len(y)//hop]`` power mel spectrogram (src/spectrogram_explainability.py:368-376)."""
    S = dsp.stft(y, n_fft, hop_length, win_length).numpy()
    power = (np.abs(S) ** 2).astype(np.float32)
    return slaney_mel_filterbank(sr, n_fft, n_mels, 0.0, fmax) @ power


def power_to_db_refmax(S: np.ndarray) -> np.ndarray:
    """librosa.power_to_db(S, ref=np.max) with defaults amin=1e-10, top_db=80 (visualisation only, :377)."""
    S = np.asarray(S, dtype=np.float32)
    log_spec = 10.0 * np.log10(np.maximum(1e-10, S)) - 10.0 * np.log10(max(1e-10, float(S.max())))
    return np.maximum(log_spec, log_spec.max() - 80.0)


def nnls_operators(A: np.ndarray) -> Tuple[np.ndarray, float]:
    """(pinv(A) float32 ``[n_freq, n_mels]``, 1 / L with L = ||A||_2^2) of the builder's NNLS; float64 linear algebra."""
    A64 = np.asarray(A, dtype=np.float64)
    P = np.linalg.pinv(A64).astype(np.float32)
    L = float(np.linalg.norm(A64, 2) ** 2)
    return P, np.float32(1.0 / L)


def nnls_builder(A: np.ndarray, B: np.ndarray, nnls_iter: int = 16) -> np.ndarray:
    """argmin-ish ``||A X - B||`` s.t. ``X >= 0`` for every column of ``B``: clipped least squares + projected gradient,
    float32 like the CUDA kernel (the order of the float32 sums differs; the parity tolerance covers it)."""
    A = np.asarray(A, dtype=np.float32)
    B = np.asarray(B, dtype=np.float32)
    P, step = nnls_operators(A)
    X = np.maximum(P @ B, np.float32(0))
    for _ in range(nnls_iter):
        R = A @ X - B
        X = np.maximum(X - step * (A.T @ R), np.float32(0))
    return X.astype(np.float32)


def mel_to_stft_builder(M: np.ndarray, sr: int, n_fft: int, nnls_iter: int = 16) -> np.ndarray:
    """librosa.feature.inverse.mel_to_stft(M, sr, n_fft, power=2.0) with the builder's NNLS: magnitude ``[n_freq, T]``.
    The reference call passes no fmax (:395-402), so the basis spans [0, sr/2] whatever the forward fmax was."""
    A = slaney_mel_filterbank(sr, n_fft, M.shape[0], 0.0, None)
    return np.sqrt(nnls_builder(A, M, nnls_iter))


def phase_uniform(seed: int, index: int, n_freq: int, n_time: int) -> np.ndarray:
    """u in [0, 1) per cell (frame t, bin k): hash_lowbias32(key(seed, index) ^ ((t * 1025 + k) * 0xC2B2AE35)) / 2^32."""
    m32 = np.uint64(0xFFFFFFFF)
    key = _lowbias32(np.array([(np.uint64(seed & 0xFFFFFFFF) * np.uint64(0x9E3779B9) + np.uint64(index) * np.uint64(0x85EBCA6B)
                                + np.uint64(0x165667B1)) & m32]))[0]
    t = np.arange(n_time, dtype=np.uint64)[None, :]
    f = np.arange(n_freq, dtype=np.uint64)[:, None]
    cell = (t * np.uint64(1025) + f) & m32
    h = _lowbias32(key ^ ((cell * np.uint64(0xC2B2AE35)) & m32))
    return (h.astype(np.float64) / 4294967296.0).astype(np.float32)


def griffinlim_builder(mag: np.ndarray, n_iter: int, seed: int, index: int, hop_length: int = 512, win_length: int = 2048,
                       momentum: float = 0.99) -> np.ndarray:
    """librosa.griffinlim(S, n_iter, hop_length, win_length, n_fft, momentum=0.99, init='random') with hashed phases."""
    mag = np.asarray(mag, dtype=np.float32)
    n_fft = 2 * (mag.shape[0] - 1)
    u = phase_uniform(seed, index, mag.shape[0], mag.shape[1])
    angles = np.empty(mag.shape, dtype=np.complex64)
    angles.real = np.cos(2 * np.pi * u.astype(np.float64)).astype(np.float32)
    angles.imag = np.sin(2 * np.pi * u.astype(np.float64)).astype(np.float32)
    eps = np.finfo(np.float32).tiny
    rebuilt = np.zeros_like(angles)
    tprev = np.zeros_like(angles)
    for _ in range(n_iter):
        tprev = rebuilt
        inverse = dsp.istft(mag * angles, hop_length, win_length)
        rebuilt = dsp.stft(inverse, n_fft, hop_length, win_length).numpy()
        angles = rebuilt - np.float32(momentum / (1 + momentum)) * tprev
        angles = (angles / (np.abs(angles) + eps)).astype(np.complex64)
    return dsp.istft(mag * angles, hop_length, win_length).numpy()


def mel_to_audio_builder(M: np.ndarray, sr: int, n_fft: int = 2048, hop_length: int = 512, win_length: int = 2048,
                         n_iter: int = 32, seed: int = 0, index: int = 0, nnls_iter: int = 16) -> np.ndarray:
    """``_invert_spectrogram`` for ``spec_type == 'mel'`` (:394-402)."""
    return griffinlim_builder(mel_to_stft_builder(M, sr, n_fft, nnls_iter), n_iter, seed, index, hop_length, win_length)


class MelOcclusionOut(NamedTuple):
    importance_map: Optional[np.ndarray]
    baseline_pred: float
    patch_importances: Optional[List[dict]]
    y: np.ndarray
    S: np.ndarray


def occlusion_map_mel(y: np.ndarray, predictor, sr: int, n_mels: int = 128, n_iter: int = 32, n_fft: int = 2048,
                      hop_length: int = 512, win_length: int = 2048, fmax: Optional[float] = None,
                      patch_time_frames: int = 1024, stride_time_frames: int = 1024, patch_freq_percent: float = 5.0,
                      stride_freq_percent: float = 5.0, occlusion_value: float = 0.0, baseline_threshold: float = 0.3,
                      seed: int = 0, nnls_iter: int = 16, windows: Optional[Sequence[int]] = None) -> MelOcclusionOut:
    """``_compute_occlusion_map`` with ``spec_type='mel'`` (:589-720): the loop is the STFT one, over the mel spectrogram;
    window i inverts with phase index i.  ``windows`` restricts the loop to a subset of the grid (tests)."""
    y = np.asarray(y, dtype=np.float32)
    S = melspectrogram(y, sr, n_fft, hop_length, win_length, n_mels, fmax)
    baseline_pred = float(predictor.predict(y, sr))
    if baseline_pred < baseline_threshold:
        return MelOcclusionOut(None, baseline_pred, None, y, S)
    n_freq, n_time = S.shape
    importance_map = np.zeros((n_freq, n_time))
    count_map = np.zeros((n_freq, n_time))
    patch_freq = max(1, int(round(patch_freq_percent / 100.0 * n_freq)))
    stride_freq = max(1, int(round(stride_freq_percent / 100.0 * n_freq)))
    positions = [(t, f) for t in range(0, max(1, n_time - patch_time_frames + 1), stride_time_frames)
                 for f in range(0, max(1, n_freq - patch_freq + 1), stride_freq)]
    patches: List[dict] = []
    for i, (t0, f0) in enumerate(positions):
        if windows is not None and i not in windows:
            continue
        t1, f1 = min(t0 + patch_time_frames, n_time), min(f0 + patch_freq, n_freq)
        S_occ = S.copy()
        S_occ[f0:f1, t0:t1] = occlusion_value
        y_occ = mel_to_audio_builder(S_occ, sr, n_fft, hop_length, win_length, n_iter, seed, i, nnls_iter)
        if len(y_occ) > len(y):
            y_occ = y_occ[: len(y)]
        elif len(y_occ) < len(y):
            y_occ = np.pad(y_occ, (0, len(y) - len(y_occ)))
        importance = baseline_pred - float(predictor.predict(y_occ, sr))
        patches.append({"t_start": t0, "t_end": t1, "f_start": f0, "f_end": f1, "importance": importance})
        importance_map[f0:f1, t0:t1] += importance
        count_map[f0:f1, t0:t1] += 1
    return MelOcclusionOut(importance_map / (count_map + 1e-8), baseline_pred, patches, y, S)


def mel_band_gains(bands: Sequence[Tuple[float, float]], sr: float, n_mels: int, attenuation: float, transition_mode: str,
                   transition_rel: float, transition_min_hz: float, transition_max_hz: float, transition_hz: float,
                   fmax: Optional[float] = None) -> np.ndarray:
    """FBP-mel (builder-defined): the band keep mask of src/dsp_band_ops.py:236-259, 576 evaluated at the CENTRE frequency of
    every mel bin (mel_frequencies(n_mels + 2)[1:-1]) instead of at the FFT bin frequencies -> float64 ``[n_bands, n_mels]``."""
    from .loops import _keep_mask
    centres = mel_frequencies(n_mels + 2, 0.0, float(sr) / 2 if fmax is None else float(fmax))[1:-1]
    out = np.empty((len(bands), n_mels))
    for b, (low, high) in enumerate(bands):
        bw = float(high - low)
        trans = float(np.clip(bw * transition_rel, transition_min_hz, transition_max_hz)) if transition_mode == "rel" else float(transition_hz)
        keep = _keep_mask(centres, low, high, trans)
        out[b] = keep + attenuation * (1.0 - keep)
    return out
