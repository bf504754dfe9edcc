"""ORACLE (test infrastructure only - never imported by the product path).

CPU restatement of the DSP primitives on the perturbation hot path.  The arithmetic lives in the
un-vendored, un-pinned third-party package ``librosa`` (call sites: src/spectrogram_explainability.py:
368-410, 601; src/dsp_band_ops.py:383-425, 566-567) and, for the classifier front-end, in
``torchaudio`` inside ``sonics``.  ``librosa`` is absent from this image, so its published algorithm
(librosa >= 0.10 ``stft``/``istft``/``magphase``/``fft_frequencies``) is restated here in float32 torch:

PARITY STATUS: the librosa numerics are *restated, unpinned* (no librosa, no reference fixtures exist);
they are cross-checked against ``torch.stft``/``torch.istft`` (tests/test_oracle_dsp.py).  The mel
front-end calls the real ``torchaudio`` (pinned to the library ``sonics`` itself uses).
"""
from __future__ import annotations

import numpy as np
import torch


def hann_periodic(win_length: int, dtype=torch.float32) -> torch.Tensor:
    """``scipy.signal.get_window('hann', n, fftbins=True)`` (librosa's window), computed in float64."""
    n = torch.arange(win_length, dtype=torch.float64)
    return (0.5 - 0.5 * torch.cos(2.0 * torch.pi * n / win_length)).to(dtype)


def stft(y, n_fft: int = 2048, hop_length: int = 512, win_length: int = 2048) -> torch.Tensor:
    """librosa.stft(y, n_fft, hop, win, window='hann', center=True, pad_mode='constant') -> complex64
    ``[1 + n_fft//2, 1 + len(y)//hop]`` (src/spectrogram_explainability.py:379-386)."""
    assert win_length == n_fft, "reference always uses win_length == n_fft"
    y = torch.as_tensor(np.asarray(y), dtype=torch.float32)
    pad = n_fft // 2
    yp = torch.nn.functional.pad(y, (pad, pad))                    # zero padding (pad_mode='constant')
    frames = yp.unfold(0, n_fft, hop_length)                        # [n_frames, n_fft]
    spec = torch.fft.rfft(frames * hann_periodic(win_length), dim=-1)
    return spec.transpose(0, 1).contiguous().to(torch.complex64)    # [n_freq, n_frames]


def window_sumsquare(n_frames: int, n_fft: int, hop_length: int, win_length: int) -> torch.Tensor:
    """librosa.filters.window_sumsquare(norm=None): overlap-added squared window, float32."""
    w2 = hann_periodic(win_length) ** 2
    out = torch.zeros(n_fft + hop_length * (n_frames - 1), dtype=torch.float32)
    for i in range(n_frames):
        out[i * hop_length: i * hop_length + n_fft] += w2
    return out


def istft(S, hop_length: int = 512, win_length: int = 2048) -> torch.Tensor:
    """librosa.istft(S, hop, win, window='hann', center=True, length=None): per-frame irfft * window,
    overlap-add, divide by the window sum-square where it exceeds ``tiny``, trim ``n_fft//2`` at both
    ends -> real ``[hop * (n_frames - 1)]`` (src/spectrogram_explainability.py:404-410).
    Computes in the real dtype matching S (complex64 -> float32, complex128 -> float64) as librosa does."""
    S = torch.as_tensor(np.asarray(S)) if not isinstance(S, torch.Tensor) else S
    n_fft = 2 * (S.shape[0] - 1)
    assert win_length == n_fft
    rdtype = torch.float64 if S.dtype == torch.complex128 else torch.float32
    n_frames = S.shape[1]
    w = hann_periodic(win_length, rdtype)
    ytmp = torch.fft.irfft(S.transpose(0, 1), n=n_fft, dim=-1).to(rdtype) * w   # [n_frames, n_fft]
    total = n_fft + hop_length * (n_frames - 1)
    # overlap-add via fold (same sums as librosa's __overlap_add loop up to addition order)
    y = torch.nn.functional.fold(
        ytmp.transpose(0, 1).unsqueeze(0), output_size=(1, total), kernel_size=(1, n_fft), stride=(1, hop_length)
    ).reshape(total)
    wss = torch.nn.functional.fold(
        (w * w).unsqueeze(1).expand(n_fft, n_frames).unsqueeze(0).contiguous(),
        output_size=(1, total), kernel_size=(1, n_fft), stride=(1, hop_length),
    ).reshape(total)
    start = n_fft // 2
    y, wss = y[start: total - start], wss[start: total - start]
    nz = wss > torch.finfo(rdtype).tiny
    y = torch.where(nz, y / torch.where(nz, wss, torch.ones_like(wss)), y)
    return y


def amplitude_to_db_refmax(mag) -> np.ndarray:
    """librosa.amplitude_to_db(np.abs(S), ref=np.max) with defaults amin=1e-5, top_db=80 (viz only)."""
    mag = np.abs(np.asarray(mag)).astype(np.float32)
    power = mag ** 2
    ref = float(mag.max()) ** 2
    log_spec = 10.0 * np.log10(np.maximum(1e-10, power)) - 10.0 * np.log10(max(1e-10, ref))
    return np.maximum(log_spec, log_spec.max() - 80.0)


def magphase(S):
    """librosa.magphase: (abs(D), D/abs(D)) with phase = 1 where abs == 0."""
    S = np.asarray(S)
    mag = np.abs(S)
    zeros = mag == 0
    phase = np.where(zeros, 1.0, S / np.where(zeros, 1.0, mag)).astype(S.dtype)
    return mag, phase


def fft_frequencies(sr: float, n_fft: int) -> np.ndarray:
    return np.fft.rfftfreq(n=n_fft, d=1.0 / sr)


def match_rms(ref: np.ndarray, x: np.ndarray, eps: float = 1e-8) -> np.ndarray:
    """src/dsp_band_ops.py:228-233."""
    r_ref = float(np.sqrt(np.mean(ref ** 2) + eps))
    r_x = float(np.sqrt(np.mean(x ** 2) + eps))
    if r_x < eps:
        return x
    return x * (r_ref / r_x)


# --------------------------------------------------------------------------------------------------
# classifier front-end (third-party torchaudio inside third-party sonics)
# --------------------------------------------------------------------------------------------------
_MEL_CACHE: dict = {}


def mel_frontend(audio: torch.Tensor, cfg, stage: str = "norm") -> torch.Tensor:
    """sonics ``FeatureExtractor``: torchaudio MelSpectrogram(power=2, HTK, norm=None, center, reflect)
    -> AmplitudeToDB('power', top_db) -> per-sample (x-mean)/(std+eps).  ``audio`` is ``[B, L]`` float32.
    The top_db clamp is applied per sample (the reference only ever calls the model with batch 1,
    src/sonics_api.py:269).  ``stage``: 'power' | 'db' | 'norm' selects the returned intermediate."""
    import torchaudio

    key = (cfg.sample_rate, cfg.n_fft, cfg.hop_length, cfg.win_length, cfg.n_mels, cfg.f_min, cfg.f_max)
    if key not in _MEL_CACHE:
        _MEL_CACHE[key] = torchaudio.transforms.MelSpectrogram(
            sample_rate=cfg.sample_rate, n_fft=cfg.n_fft, win_length=cfg.win_length, hop_length=cfg.hop_length,
            f_min=cfg.f_min, f_max=cfg.f_max, n_mels=cfg.n_mels, power=2.0,
        )
    mel = _MEL_CACHE[key](audio.float())                                    # [B, n_mels, n_frames]
    if stage == "power":
        return mel
    to_db = torchaudio.transforms.AmplitudeToDB(stype="power", top_db=cfg.top_db)
    db = torch.stack([to_db(m.unsqueeze(0)).squeeze(0) for m in mel])        # per-sample top_db clamp
    if stage == "db":
        return db
    mean = db.mean((1, 2), keepdim=True)
    std = db.std((1, 2), keepdim=True, unbiased=cfg.std_unbiased)
    return (db - mean) / (std + cfg.norm_eps)


def mel_filterbank(cfg) -> torch.Tensor:
    """torchaudio.functional.melscale_fbanks(htk, norm=None): ``[n_freq, n_mels]`` float32."""
    import torchaudio

    return torchaudio.functional.melscale_fbanks(
        n_freqs=cfg.n_fft // 2 + 1, f_min=cfg.f_min, f_max=cfg.f_max, n_mels=cfg.n_mels,
        sample_rate=cfg.sample_rate, norm=None, mel_scale="htk",
    )
