"""Track loading for the drop-in entry points (the step before the hot path).

The reference calls ``librosa.load(path, sr=sr, duration=duration, mono=True)``
(src/spectrogram_explainability.py:601, src/dsp_band_ops.py:679) and writes with ``soundfile.write(path, y, sr)``
(:494, dsp_band_ops.py:636), i.e. 16-bit PCM WAV.

Decoding: whatever the host offers, in the reference's own order of preference - ``librosa`` (then the samples are the
reference's, bit for bit, resampler included), ``soundfile``, ``torchaudio`` - and a built-in WAV / ``.npy`` reader
(scipy) otherwise.  None of the three libraries and no mp3 decoder is part of this image, so here the built-in reader is
what runs; a file nothing can decode raises ``AudioDecodeError`` and the dataset drivers log it and move on.
Resampling without librosa: polyphase filtering (``scipy.signal.resample_poly`` on the host, or the engine's CUDA polyphase
kernel through ``resample_on_device``); librosa's default is ``soxr_hq``, so sample values match the reference only when
the file is already at ``sr`` or librosa itself is importable - INTEGRATION.md states this.
"""
from __future__ import annotations

from math import gcd
from pathlib import Path
from typing import Callable, Optional, Tuple

import numpy as np


class AudioDecodeError(RuntimeError):
    """The file could not be decoded by any available backend."""


def _mono(data: np.ndarray, channels_last: bool) -> np.ndarray:
    if data.ndim == 2:
        data = data.mean(axis=1 if channels_last else 0)
    return np.asarray(data, dtype=np.float32)


def _read_native(path: Path) -> Tuple[np.ndarray, int]:
    """(mono float32 samples, native sample rate) through the first backend that can read the file."""
    suffix = path.suffix.lower()
    errors = []
    try:                                        # 1. soundfile: what librosa itself uses for wav / flac / ogg
        import soundfile as sf

        data, native_sr = sf.read(str(path), dtype="float32", always_2d=False)
        return _mono(data, True), int(native_sr)
    except ImportError:
        pass
    except Exception as e:                      # unreadable for soundfile (mp3 on old libsndfile, ...): try the next backend
        errors.append(f"soundfile: {e}")
    try:                                        # 2. torchaudio (ffmpeg / sox backends decode mp3 where they are installed)
        import torchaudio

        wav, native_sr = torchaudio.load(str(path))
        return _mono(wav.numpy(), False), int(native_sr)
    except ImportError:
        pass
    except Exception as e:
        errors.append(f"torchaudio: {e}")
    if suffix == ".wav":                        # 3. built-in WAV reader
        from scipy.io import wavfile

        try:
            native_sr, data = wavfile.read(str(path))
        except Exception as e:
            raise AudioDecodeError(f"cannot decode '{path.name}': {e}") from e
        if data.dtype.kind == "i":
            data = data.astype(np.float32) / float(np.iinfo(data.dtype).max + 1)
        elif data.dtype.kind == "u":
            data = (data.astype(np.float32) - 128.0) / 128.0
        return _mono(np.asarray(data), True), int(native_sr)
    raise AudioDecodeError(f"cannot decode '{path.name}': no decoder for '{suffix}' on this host "
                           f"(librosa / soundfile / a working torchaudio backend absent{'; ' + '; '.join(errors) if errors else ''})")


def load_audio(path, sr: Optional[int] = 22050, duration: Optional[float] = None, mono: bool = True,
               resample: Optional[Callable[[np.ndarray, int, int], np.ndarray]] = None) -> Tuple[np.ndarray, int]:
    """``librosa.load(path, sr=sr, duration=duration, mono=True)``.  ``resample(data, native_sr, sr)`` overrides the host
    polyphase resampler (the explainers pass the engine's CUDA resampler)."""
    path = Path(path)
    if path.suffix.lower() == ".npy":
        data = np.load(path)
        data = _mono(data, data.ndim == 2 and data.shape[1] <= 8)
        if duration is not None and sr:
            data = data[: int(round(duration * sr))]
        return np.ascontiguousarray(data), int(sr) if sr else 0
    try:
        import librosa                          # the reference's own loader: identical samples, soxr_hq resampling

        y, out_sr = librosa.load(str(path), sr=sr, duration=duration, mono=mono)
        return np.ascontiguousarray(y, dtype=np.float32), int(out_sr)
    except ImportError:
        pass
    data, native_sr = _read_native(path)
    if duration is not None:
        data = data[: int(round(duration * native_sr))]
    if sr is not None and native_sr != sr:
        data = resample(data, native_sr, int(sr)) if resample is not None else resample_poly_host(data, native_sr, int(sr))
    return np.ascontiguousarray(data, dtype=np.float32), int(sr if sr is not None else native_sr)


def polyphase_filter(up: int, down: int, half_width: int = 16, beta: float = 8.6) -> np.ndarray:
    """Kaiser-windowed sinc low-pass for rational resampling by up / down, float64, unit DC gain (the resamplers apply the
    x up interpolation gain, like scipy.signal.resample_poly does with a user filter); the design of resample_poly's default
    ('kaiser', 5.0) with a longer, sharper window.  Odd length 2 * half_width * max(up, down) + 1, shared by the host and
    CUDA resamplers."""
    m = max(up, down)
    n = 2 * half_width * m + 1
    t = np.arange(n, dtype=np.float64) - (n - 1) / 2
    h = np.sinc(t / m) * np.kaiser(n, beta)
    return h / h.sum()


def resample_poly_host(data: np.ndarray, native_sr: int, sr: int) -> np.ndarray:
    """Rational polyphase resampling on the host with ``polyphase_filter`` (float64 accumulate): y[n] = sum_k h[n down - k up] x[k],
    centred (no group delay), output length ceil(len * up / down)."""
    g = gcd(int(sr), int(native_sr))
    up, down = int(sr) // g, int(native_sr) // g
    from scipy.signal import resample_poly

    return resample_poly(np.asarray(data, dtype=np.float64), up, down, window=polyphase_filter(up, down)).astype(np.float32)


def write_wav(path, data: np.ndarray, sr: int) -> None:
    """``soundfile.write(path, y, sr)`` for a float array: 16-bit PCM WAV, the value mapping of libsndfile's float -> short
    conversion (scale by 2^15, round to nearest, clip), src/spectrogram_explainability.py:494, src/dsp_band_ops.py:636."""
    try:
        import soundfile as sf

        sf.write(str(path), np.asarray(data), int(sr))
        return
    except ImportError:
        pass
    from scipy.io import wavfile

    x = np.asarray(data, dtype=np.float64)
    pcm = np.clip(np.rint(x * 32768.0), -32768, 32767).astype(np.int16)
    wavfile.write(str(path), int(sr), pcm)
