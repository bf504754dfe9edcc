"""Seeded synthetic tracks standing in for the reference dataset folders
(configs/Spec_occlusion_configs/spectrogram_explainability.yaml:6-11; the real mp3s are git-ignored).

Five generator families; each returns float32 mono, peak-normalised to 0.5, and
can also return its four additive stems (they sum to the un-normalised mix) for the stem-mask sweep.
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np

FAMILIES = ("REAL", "SUNO", "SUNO_PRO", "UDIO", "ElevenLabs")
_SEED_BASE = {"REAL": 1000, "SUNO": 2000, "SUNO_PRO": 3000, "UDIO": 4000, "ElevenLabs": 5000}


def _pink(rng: np.random.Generator, n: int) -> np.ndarray:
    spec = np.fft.rfft(rng.standard_normal(n))
    f = np.arange(spec.shape[0], dtype=np.float64)
    f[0] = 1.0
    x = np.fft.irfft(spec / np.sqrt(f), n)
    return x / (np.abs(x).max() + 1e-12)


def _lowpass(x: np.ndarray, sr: int, cutoff: float) -> np.ndarray:
    spec = np.fft.rfft(x)
    spec[np.fft.rfftfreq(x.shape[0], 1.0 / sr) > cutoff] = 0.0
    return np.fft.irfft(spec, x.shape[0])


def _harmonic(rng: np.random.Generator, t: np.ndarray, sr: int) -> np.ndarray:
    f0 = rng.uniform(110.0, 440.0)
    vib = 1.0 + 0.01 * np.sin(2 * np.pi * rng.uniform(4.0, 6.0) * t)
    phase = 2 * np.pi * np.cumsum(f0 * vib) / sr
    x = np.zeros_like(t)
    for h in range(1, 9):
        x += np.sin(h * phase + rng.uniform(0, 2 * np.pi)) / h
    return x * (0.6 + 0.4 * np.sin(2 * np.pi * 2.0 * t))


def synth_stems(family: str, index: int = 0, sr: int = 16000, duration: float = 120.0) -> Dict[str, np.ndarray]:
    """Four float64 stems of one synthetic track (they sum to the un-normalised mix)."""
    if family not in _SEED_BASE:
        raise ValueError(f"unknown family {family!r}; expected one of {FAMILIES}")
    rng = np.random.default_rng(_SEED_BASE[family] + index)
    n = int(round(sr * duration))
    t = np.arange(n, dtype=np.float64) / sr
    zeros = np.zeros(n)
    if family in ("REAL", "SUNO", "SUNO_PRO"):
        a = 0.3 * _pink(rng, n)
        b = 0.5 * _harmonic(rng, t, sr)
        c, d = zeros.copy(), zeros.copy()
        if family != "REAL":
            a, b = _lowpass(a, sr, 5000.0), _lowpass(b, sr, 5000.0)
            c = 0.02 * np.sin(2 * np.pi * abs(12000.0 - sr * round(12000.0 / sr)) * t)  # 12 kHz tone folded
        if family == "SUNO_PRO":
            d = 10 ** (-50 / 20) * rng.standard_normal(n)
    elif family == "UDIO":
        fc, idx = rng.uniform(220.0, 880.0), rng.uniform(2.0, 6.0)
        gate = (np.sin(2 * np.pi * 4.0 * t) > 0).astype(np.float64)
        a = 0.5 * np.sin(2 * np.pi * fc * t + idx * np.sin(2 * np.pi * fc * 0.5 * t)) * gate
        b = 0.2 * np.sin(2 * np.pi * 2 * fc * t + 0.5 * idx * np.sin(2 * np.pi * fc * t)) * gate
        c = 0.05 * _pink(rng, n)
        d = 0.1 * np.sin(2 * np.pi * 55.0 * t)
    else:  # ElevenLabs: glottal pulse train through three fixed formant resonators + silence gaps
        f0 = rng.uniform(90.0, 180.0)
        pulses = np.zeros(n)
        pulses[(np.arange(0, n, sr / f0)).astype(np.int64)] = 1.0
        gaps = (np.sin(2 * np.pi * 0.25 * t + rng.uniform(0, 2 * np.pi)) > -0.6).astype(np.float64)
        spec = np.fft.rfft(pulses * gaps)
        fr = np.fft.rfftfreq(n, 1.0 / sr)
        stems = []
        for fc, bw in ((700.0, 110.0), (1220.0, 150.0), (2600.0, 250.0)):
            stems.append(np.fft.irfft(spec / (1.0 + ((fr - fc) / bw) ** 2), n))
        a, b, c = (s / (np.abs(s).max() + 1e-12) * g for s, g in zip(stems, (0.5, 0.3, 0.15)))
        d = 0.01 * _pink(rng, n)
    return {"stem0": a, "stem1": b, "stem2": c, "stem3": d}


def synth_track(
    family: str, index: int = 0, sr: int = 16000, duration: float = 120.0, with_stems: bool = False
) -> Tuple[np.ndarray, Dict[str, np.ndarray]] | np.ndarray:
    """float32 mono track of ``family``; peak 0.5. With ``with_stems`` also the scaled float32 stems."""
    stems = synth_stems(family, index, sr, duration)
    mix = sum(stems.values())
    g = 0.5 / (np.abs(mix).max() + 1e-12)
    y = (mix * g).astype(np.float32)
    if with_stems:
        return y, {k: (v * g).astype(np.float32) for k, v in stems.items()}
    return y
