"""Track loading for the drop-in entry points (the step before the hot path).

The reference calls ``librosa.load(path, sr=sr, duration=duration, mono=True)``
(src/spectrogram_explainability.py:601, src/dsp_band_ops.py:679).  librosa and an mp3 decoder are not part of
this image, so this loader handles WAV (scipy) and ``.npy`` arrays, down-mixes to mono by channel mean and
resamples with a polyphase filter.  (librosa's default resampler is ``soxr_hq``; sample values therefore match
the reference only when the file is already at ``sr``.)
"""
from __future__ import annotations

from math import gcd
from pathlib import Path
from typing import Optional, Tuple

import numpy as np


def load_audio(path, sr: Optional[int] = 22050, duration: Optional[float] = None, mono: bool = True) -> Tuple[np.ndarray, int]:
    path = Path(path)
    suffix = path.suffix.lower()
    if suffix == ".npy":
        data = np.load(path)
        native_sr = sr
    elif suffix == ".wav":
        from scipy.io import wavfile

        native_sr, data = wavfile.read(str(path))
        if data.dtype.kind == "i":
            data = data.astype(np.float32) / float(np.iinfo(data.dtype).max + 1)
        elif data.dtype.kind == "u":
            data = (data.astype(np.float32) - 128.0) / 128.0
        else:
            data = data.astype(np.float32)
    else:
        raise RuntimeError(f"cannot decode '{path.name}': only .wav and .npy are supported in this build (no mp3 decoder)")
    if data.ndim == 2 and mono:
        data = data.mean(axis=1 if data.shape[1] <= 8 else 0)
    data = np.asarray(data, dtype=np.float32)
    if duration is not None and native_sr:
        data = data[: int(round(duration * native_sr))]
    if sr is not None and native_sr is not None and native_sr != sr:
        from scipy.signal import resample_poly

        g = gcd(int(sr), int(native_sr))
        data = resample_poly(data.astype(np.float64), int(sr) // g, int(native_sr) // g).astype(np.float32)
    return np.ascontiguousarray(data), int(sr if sr is not None else native_sr)


def write_wav(path, data: np.ndarray, sr: int) -> None:
    """float32 WAV writer standing in for ``soundfile.write`` (src/spectrogram_explainability.py:494)."""
    from scipy.io import wavfile

    wavfile.write(str(path), int(sr), np.asarray(data, dtype=np.float32))
