"""Python handle on the native engine (``b200x_engine_*`` in include/b200xai.h).

Thin: numpy host buffers in, numpy out; all arithmetic happens in libb200xai.so on the GPU.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np

from . import _lib
from .weights import ALPHA_120S, SpecTTTraConfig, random_state_dict

RANK_ABS_DESC, RANK_ABS_ASC, RANK_DESC, RANK_ASC = 0, 1, 2, 3


def _ptr(a: np.ndarray):
    return C.c_void_p(a.ctypes.data)


def _pinned(shape, dtype) -> np.ndarray:
    """Page-locked host array (torch's caching host allocator) so that device->host copies run at PCIe speed without a
    driver staging copy; the numpy view keeps the allocation alive."""
    try:
        import torch

        return torch.empty(shape, dtype=getattr(torch, np.dtype(dtype).name), pin_memory=True).numpy()
    except Exception:                      # no torch / no CUDA runtime in this interpreter: plain pageable memory
        return np.empty(shape, dtype)


def _c_config(cfg: SpecTTTraConfig) -> _lib.ModelConfig:
    return _lib.ModelConfig(
        sample_rate=cfg.sample_rate, n_fft=cfg.n_fft, hop_length=cfg.hop_length, n_mels=cfg.n_mels,
        f_min=cfg.f_min, f_max=cfg.f_max, top_db=cfg.top_db, amin=cfg.amin, norm_eps=cfg.norm_eps,
        std_unbiased=int(cfg.std_unbiased), input_spec_dim=cfg.input_spec_dim, input_temp_dim=cfg.input_temp_dim,
        t_clip=cfg.t_clip, f_clip=cfg.f_clip, embed_dim=cfg.embed_dim, num_heads=cfg.num_heads,
        num_layers=cfg.num_layers, mlp_hidden=cfg.mlp_hidden, pre_norm=int(cfg.pre_norm),
        pe_learnable=int(cfg.pe_learnable), qkv_bias=int(cfg.qkv_bias), final_norm=int(cfg.final_norm),
        tokenizer_ln_eps=cfg.tokenizer_ln_eps, block_ln_eps=cfg.block_ln_eps,
    )


class Engine:
    """One engine per GPU / process.  ``state_dict``: sonics-named float32 arrays (see weights.py)."""

    def __init__(self, cfg: SpecTTTraConfig = ALPHA_120S, state_dict: Optional[Dict[str, np.ndarray]] = None,
                 copies_per_chunk: int = 128, max_samples: int = 120 * 16000, device: int = 0):
        self.lib = _lib.load()
        self.cfg = cfg
        self.device = device
        self.copies_per_chunk = copies_per_chunk
        self.max_samples = int(max_samples)
        _lib.check(self.lib.b200x_set_device(device), "set_device")
        h = C.c_void_p()
        cc = _c_config(cfg)
        _lib.check(self.lib.b200x_engine_create(C.byref(cc), copies_per_chunk, self.max_samples, C.byref(h)), "engine_create")
        self._h = h
        self.n_samples = 0
        self.load_state_dict(state_dict if state_dict is not None else random_state_dict(cfg, 0))

    # ------------------------------------------------------------------ weights
    def load_state_dict(self, sd: Dict[str, np.ndarray]) -> None:
        for name, v in sd.items():
            if name.startswith("ft_extractor."):
                continue                       # the mel front-end has no learned parameters
            a = np.ascontiguousarray(np.asarray(v, dtype=np.float32))
            _lib.check(self.lib.b200x_engine_set_param(self._h, name.encode(), _ptr(a), a.size), f"set_param({name})")
        _lib.check(self.lib.b200x_engine_finalize(self._h), "finalize")

    def close(self) -> None:
        if getattr(self, "_h", None):
            self.lib.b200x_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ classifier
    def predict(self, waves: np.ndarray, return_logits: bool = False):
        """Fake-probability of each equal-length wave (``[L]`` or ``[count, L]``, float32 after ``.float()``)."""
        w = np.ascontiguousarray(np.asarray(waves, dtype=np.float32))
        single = w.ndim == 1
        if single:
            w = w[None]
        prob = np.empty(w.shape[0], np.float32)
        logit = np.empty(w.shape[0], np.float32)
        _lib.check(self.lib.b200x_engine_predict(self._h, _ptr(w), w.shape[1], w.shape[0], 0, _ptr(prob), _ptr(logit)), "predict")
        if return_logits:
            return (prob[0], logit[0]) if single else (prob, logit)
        return prob[0] if single else prob

    def predict_track(self) -> float:
        """Fake-probability of the track loaded with ``set_track`` (device-resident samples, no second upload)."""
        prob = np.empty(1, np.float32)
        _lib.check(self.lib.b200x_engine_predict_track(self._h, _ptr(prob), None), "predict_track")
        return prob[0]

    # ------------------------------------------------------------------ track state
    def set_track(self, wave: np.ndarray) -> None:
        w = np.ascontiguousarray(np.asarray(wave, dtype=np.float32))
        _lib.check(self.lib.b200x_engine_set_track(self._h, _ptr(w), w.shape[0], 0), "set_track")
        self.n_samples = int(w.shape[0])

    def track_shape(self):
        f, t = C.c_int32(), C.c_int32()
        _lib.check(self.lib.b200x_engine_track_shape(self._h, C.byref(f), C.byref(t)), "track_shape")
        return int(f.value), int(t.value)

    def spectrogram(self) -> np.ndarray:
        f, t = self.track_shape()
        out = np.empty((f, t), np.complex64)
        _lib.check(self.lib.b200x_engine_get_spectrogram(self._h, _ptr(out)), "get_spectrogram")
        return out

    # ------------------------------------------------------------------ sweeps
    def occlusion_sweep(self, windows: np.ndarray, occlusion_value: float = 0.0, with_baseline: bool = False):
        """Fake-probability of every occluded copy; ``with_baseline=True`` also evaluates the track itself in the same
        device pass and returns ``(prob, baseline)`` with ``baseline == predict_track()`` bit for bit."""
        w = np.ascontiguousarray(np.asarray(windows, dtype=np.int32)).reshape(-1, 4)
        prob = np.empty(w.shape[0], np.float32)
        if with_baseline:
            base = np.empty(1, np.float32)
            _lib.check(self.lib.b200x_engine_occlusion_sweep_base(self._h, _ptr(w), w.shape[0], float(occlusion_value), 0, _ptr(prob),
                                                                  _ptr(base)), "occlusion_sweep_base")
            return prob, base[0]
        _lib.check(self.lib.b200x_engine_occlusion_sweep(self._h, _ptr(w), w.shape[0], float(occlusion_value), 0, _ptr(prob)),
                   "occlusion_sweep")
        return prob

    def fbp_sweep(self, gains: np.ndarray, normalize_loudness: bool) -> np.ndarray:
        g = np.ascontiguousarray(np.asarray(gains, dtype=np.float32))
        n_freq, _ = self.track_shape()
        if g.ndim != 2 or g.shape[1] != n_freq:
            raise ValueError(f"gains must be [n_bands, {n_freq}], got {g.shape}")
        prob = np.empty(g.shape[0], np.float32)
        _lib.check(self.lib.b200x_engine_fbp_sweep(self._h, _ptr(g), g.shape[0], int(bool(normalize_loudness)), 0, _ptr(prob)),
                   "fbp_sweep")
        return prob

    def fbp_sweep_tracks(self, waves: np.ndarray, gains: np.ndarray, normalize_loudness: bool):
        """FBP of a batch of equal-length tracks in shared launches: ``(baseline [n_tracks], prob [n_tracks, n_bands])``."""
        w = np.ascontiguousarray(np.asarray(waves, dtype=np.float32))
        g = np.ascontiguousarray(np.asarray(gains, dtype=np.float32))
        if w.ndim != 2 or g.ndim != 2 or g.shape[1] != 1025:
            raise ValueError(f"waves [n_tracks, L] / gains [n_bands, 1025] expected, got {w.shape} / {g.shape}")
        base = np.empty(w.shape[0], np.float32)
        prob = np.empty((w.shape[0], g.shape[0]), np.float32)
        _lib.check(self.lib.b200x_engine_fbp_sweep_tracks(self._h, _ptr(w), w.shape[0], w.shape[1], _ptr(g), g.shape[0],
                                                          int(bool(normalize_loudness)), _ptr(base), _ptr(prob)), "fbp_sweep_tracks")
        self.n_samples = int(w.shape[1])
        return base, prob

    def stem_sweep(self, stems: np.ndarray, masks: np.ndarray) -> np.ndarray:
        s = np.ascontiguousarray(np.asarray(stems, dtype=np.float32))
        m = np.ascontiguousarray(np.asarray(masks) != 0).astype(np.uint8)
        if s.ndim != 2 or m.ndim != 2 or m.shape[1] != s.shape[0]:
            raise ValueError(f"stems [n_stems, L] / masks [n, n_stems] expected, got {s.shape} / {m.shape}")
        prob = np.empty(m.shape[0], np.float32)
        _lib.check(self.lib.b200x_engine_stem_sweep(self._h, _ptr(s), s.shape[0], s.shape[1], _ptr(m), m.shape[0], 0, _ptr(prob)),
                   "stem_sweep")
        return prob

    def window_audio(self, windows: np.ndarray) -> list:
        """Patch-only iSTFT audio of each window, sliced to the window's own time span (list of float32 arrays)."""
        w = np.ascontiguousarray(np.asarray(windows, dtype=np.int32)).reshape(-1, 4)
        if len(w) == 0:
            return []
        hop = self.cfg.hop_length
        stride = int(max(1, int((w[:, 1] - w[:, 0]).max()) * hop))
        out = _pinned((w.shape[0], stride), np.float32)
        lens = np.zeros(w.shape[0], np.int64)
        _lib.check(self.lib.b200x_engine_window_audio(self._h, _ptr(w), w.shape[0], _ptr(out), stride, _ptr(lens)), "window_audio")
        return [out[i, : int(lens[i])] for i in range(w.shape[0])]

    def occluded_audio(self, windows: np.ndarray, occlusion_value: float = 0.0) -> np.ndarray:
        w = np.ascontiguousarray(np.asarray(windows, dtype=np.int32)).reshape(-1, 4)
        _, t = self.track_shape()
        out = np.empty((w.shape[0], self.cfg.hop_length * (t - 1)), np.float32)
        _lib.check(self.lib.b200x_engine_occluded_audio(self._h, _ptr(w), w.shape[0], float(occlusion_value), _ptr(out)), "occluded_audio")
        return out

    def band_audio(self, gains: np.ndarray) -> np.ndarray:
        g = np.ascontiguousarray(np.asarray(gains, dtype=np.float32))
        _, t = self.track_shape()
        out = np.empty((g.shape[0], self.cfg.hop_length * (t - 1)), np.float32)
        _lib.check(self.lib.b200x_engine_band_audio(self._h, _ptr(g), g.shape[0], _ptr(out)), "band_audio")
        return out

    # ------------------------------------------------------------------ track loader front
    def resample(self, data: np.ndarray, native_sr: int, sr: int) -> np.ndarray:
        """Polyphase resampling on the device (same filter and definition as ``audio_io.resample_poly_host``); the
        ``resample=`` hook of ``audio_io.load_audio``."""
        from math import gcd

        from .audio_io import polyphase_filter
        g = gcd(int(sr), int(native_sr))
        up, down = int(sr) // g, int(native_sr) // g
        x = np.ascontiguousarray(np.asarray(data, dtype=np.float32))
        h = np.ascontiguousarray(polyphase_filter(up, down), dtype=np.float64)
        n_out = -(-x.shape[0] * up // down)
        y = np.empty(n_out, np.float32)
        _lib.check(self.lib.b200x_engine_resample(self._h, _ptr(x), x.shape[0], up, down, _ptr(h), h.shape[0], _ptr(y), n_out), "resample")
        return y

    # ------------------------------------------------------------------ mel-domain variant (spec_type: mel)
    def set_mel_basis(self, basis: np.ndarray, pinv: np.ndarray, step: float) -> None:
        """Upload the mel filterbank ``[n_mels, 1025]``, its pseudo-inverse ``[1025, n_mels]`` and the NNLS step (mel_host.py)."""
        b = np.ascontiguousarray(np.asarray(basis, dtype=np.float32))
        p = np.ascontiguousarray(np.asarray(pinv, dtype=np.float32))
        if b.ndim != 2 or b.shape[1] != 1025 or p.shape != (1025, b.shape[0]):
            raise ValueError(f"basis [n_mels, 1025] / pinv [1025, n_mels] expected, got {b.shape} / {p.shape}")
        _lib.check(self.lib.b200x_engine_set_mel_basis(self._h, b.shape[0], _ptr(b), _ptr(p), float(step)), "set_mel_basis")
        self.n_mels = int(b.shape[0])

    def mel_spectrogram(self) -> np.ndarray:
        """Power mel spectrogram float32 ``[n_mels, n_time]`` of the current track (librosa.feature.melspectrogram)."""
        _, t = self.track_shape()
        out = np.empty((self.n_mels, t), np.float32)
        _lib.check(self.lib.b200x_engine_mel_spectrogram(self._h, _ptr(out)), "mel_spectrogram")
        return out

    MASK_NONE, MASK_OCCLUDE, MASK_BAND_GAIN, MASK_KEEP_ONLY = 0, 1, 2, 3

    def mel_sweep(self, mode: int, items: Optional[np.ndarray], n_iter: int, nnls_iter: int = 16, seed: int = 0, first_index: int = 0,
                  occlusion_value: float = 0.0, momentum: float = 0.99, want_prob: bool = True, want_audio: bool = False):
        """Mel-variant sweep: every perturbed mel spectrogram -> NNLS -> Griffin-Lim -> (probability, audio).  ``items``:
        windows ``[n, 4]`` (t0, t1, mel0, mel1) for MASK_OCCLUDE / MASK_KEEP_ONLY, gains ``[n, n_mels]`` for MASK_BAND_GAIN,
        ``None`` (one unperturbed copy) for MASK_NONE."""
        wins = gains = None
        if mode in (self.MASK_OCCLUDE, self.MASK_KEEP_ONLY):
            wins = np.ascontiguousarray(np.asarray(items, dtype=np.int32)).reshape(-1, 4)
            n = wins.shape[0]
        elif mode == self.MASK_BAND_GAIN:
            gains = np.ascontiguousarray(np.asarray(items, dtype=np.float32))
            if gains.ndim != 2 or gains.shape[1] != self.n_mels:
                raise ValueError(f"gains must be [n, {self.n_mels}], got {gains.shape}")
            n = gains.shape[0]
        else:
            n = 1
        _, t = self.track_shape()
        prob = np.empty(n, np.float32) if want_prob else None
        audio = np.empty((n, self.cfg.hop_length * (t - 1)), np.float32) if want_audio else None
        _lib.check(self.lib.b200x_engine_mel_sweep(self._h, int(mode), _ptr(wins) if wins is not None else None,
                                                   _ptr(gains) if gains is not None else None, n, float(occlusion_value), int(n_iter),
                                                   int(nnls_iter), int(seed) & 0xFFFFFFFF, int(first_index), float(momentum),
                                                   _ptr(prob) if prob is not None else None, _ptr(audio) if audio is not None else None),
                   "mel_sweep")
        if want_prob and want_audio:
            return prob, audio
        return prob if want_prob else audio

    def saliency_map_shape(self, windows: np.ndarray, delta: np.ndarray, n_freq: int, n_time: int) -> np.ndarray:
        """``saliency_map`` over an arbitrary ``[n_freq, n_time]`` grid (mel spectrogram rows)."""
        w = np.ascontiguousarray(np.asarray(windows, dtype=np.int32)).reshape(-1, 4)
        d = np.ascontiguousarray(np.asarray(delta, dtype=np.float64))
        out = _pinned((n_freq, n_time), np.float64)
        _lib.check(self.lib.b200x_engine_saliency_map_shape(self._h, _ptr(w), _ptr(d), w.shape[0], int(n_freq), int(n_time), _ptr(out)),
                   "saliency_map_shape")
        return out

    # ------------------------------------------------------------------ RISE (random keep masks generated on the device)
    def rise_sweep(self, n_masks: int, seed: int, keep_probability: float, first_mask: int = 0) -> np.ndarray:
        prob = np.empty(int(n_masks), np.float32)
        _lib.check(self.lib.b200x_engine_rise_sweep(self._h, int(first_mask), int(n_masks), int(seed) & 0xFFFFFFFF,
                                                    float(keep_probability), 0, _ptr(prob)), "rise_sweep")
        return prob

    def rise_audio(self, n_masks: int, seed: int, keep_probability: float, first_mask: int = 0) -> np.ndarray:
        _, t = self.track_shape()
        out = np.empty((int(n_masks), self.cfg.hop_length * (t - 1)), np.float32)
        _lib.check(self.lib.b200x_engine_rise_audio(self._h, int(first_mask), int(n_masks), int(seed) & 0xFFFFFFFF,
                                                    float(keep_probability), _ptr(out)), "rise_audio")
        return out

    def rise_map(self, predictions: np.ndarray, seed: int, keep_probability: float) -> np.ndarray:
        """``sum_i mask_i * pred_i / (n_masks * p + 1e-8)`` (float64 ``[n_freq, n_time]``, before the min-max scaling)."""
        pr = np.ascontiguousarray(np.asarray(predictions, dtype=np.float64))
        f, t = self.track_shape()
        out = _pinned((f, t), np.float64)
        _lib.check(self.lib.b200x_engine_rise_map(self._h, _ptr(pr), pr.shape[0], int(seed) & 0xFFFFFFFF, float(keep_probability),
                                                  _ptr(out)), "rise_map")
        return out

    # ------------------------------------------------------------------ reductions
    def saliency_map(self, windows: np.ndarray, delta: np.ndarray) -> np.ndarray:
        w = np.ascontiguousarray(np.asarray(windows, dtype=np.int32)).reshape(-1, 4)
        d = np.ascontiguousarray(np.asarray(delta, dtype=np.float64))
        f, t = self.track_shape()
        out = _pinned((f, t), np.float64)
        _lib.check(self.lib.b200x_engine_saliency_map(self._h, _ptr(w), _ptr(d), w.shape[0], _ptr(out)), "saliency_map")
        return out

    def band_map(self, band_rows: np.ndarray, delta: np.ndarray) -> np.ndarray:
        r = np.ascontiguousarray(np.asarray(band_rows, dtype=np.int32)).reshape(-1, 2)
        d = np.ascontiguousarray(np.asarray(delta, dtype=np.float64))
        f, t = self.track_shape()
        out = _pinned((f, t), np.float64)
        _lib.check(self.lib.b200x_engine_band_map(self._h, _ptr(r), _ptr(d), r.shape[0], _ptr(out)), "band_map")
        return out

    def rank(self, values: np.ndarray, mode: int) -> np.ndarray:
        v = np.ascontiguousarray(np.asarray(values, dtype=np.float64))
        out = np.empty(v.shape[0], np.int32)
        _lib.check(self.lib.b200x_engine_rank(self._h, _ptr(v), v.shape[0], mode, _ptr(out)), "rank")
        return out

    # ------------------------------------------------------------------ introspection
    def debug_buffer(self, name: str):
        p, n = C.c_void_p(), C.c_int64()
        _lib.check(self.lib.b200x_engine_debug_buffer(self._h, name.encode(), C.byref(p), C.byref(n)), "debug_buffer")
        return p.value, int(n.value)

    def set_trace(self, device_ptr: Optional[int]) -> None:
        _lib.check(self.lib.b200x_engine_set_trace(self._h, C.c_void_p(device_ptr or 0)), "set_trace")

    KERNEL_CLASSES = ("istft", "mel", "resize", "gemm", "attention", "layernorm", "head", "other")

    def set_alternate(self, enable: bool) -> None:
        """Alternate the row / tile traversal direction between consecutive kernels of the forward (L2 reuse; default on)."""
        _lib.check(self.lib.b200x_engine_set_alternate(self._h, int(enable)), "set_alternate")

    def set_fused_layernorm(self, enable: bool) -> None:
        """LayerNorm as a tail of the residual GEMM before it (default) or as a separate pass; results are bit-identical."""
        _lib.check(self.lib.b200x_engine_set_fused_layernorm(self._h, int(enable)), "set_fused_layernorm")

    def set_graphs(self, enable: bool) -> None:
        """Replay the per-chunk classifier forward from a CUDA graph (default) or launch kernel by kernel."""
        _lib.check(self.lib.b200x_engine_set_graphs(self._h, int(enable)), "set_graphs")

    def set_timing(self, enable: bool) -> None:
        _lib.check(self.lib.b200x_engine_set_timing(self._h, int(enable)), "set_timing")

    def get_timing(self):
        """{class: (total_ms, launches)} from CUDA events on the engine stream since set_timing / the last call."""
        ms = np.zeros(8, np.float64)
        n = np.zeros(8, np.int64)
        _lib.check(self.lib.b200x_engine_get_timing(self._h, _ptr(ms), _ptr(n)), "get_timing")
        return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(self.KERNEL_CLASSES)}

    @property
    def launch_count(self) -> int:
        return int(self.lib.b200x_engine_launch_count(self._h))

    @property
    def stream(self) -> int:
        return int(self.lib.b200x_engine_stream(self._h) or 0)

    def synchronize(self) -> None:
        _lib.check(self.lib.b200x_engine_synchronize(self._h), "synchronize")
