"""ORACLE (test infrastructure only - never imported by the product path).

Serial CPU restatement of the reference's two perturbation loops and their reductions, driven through
the duck-typed ``predictor.predict(wave, sr) -> float`` exactly as the reference drives them:

  * occlusion_map ........... src/spectrogram_explainability.py:589-720 (``_compute_occlusion_map``)
  * top_window_groups ....... src/spectrogram_explainability.py:413-587
  * fbp_component ........... src/dsp_band_ops.py:529-666 (``_compute_component_importance``)
  * stem_mask_probs ......... src/lime_explainer.py:283-301 (``predict_fn_unified``)
  * rise_map / rise_keep_mask  src/spectrogram_explainability.py:722-806 (``_compute_rise_map``) with the build's counter-based
                              mask bits in place of the reference's UNSEEDED ``np.random.rand`` (no reference parity can exist)

PARITY STATUS: the loop / indexing / ordering logic here is *pinned*: tests/golden/ref_loops_*.npz were
produced by running the reference's own functions (imported from /root/reference with its missing
third-party imports stubbed by oracle.dsp) - see oracle/make_golden_from_reference.py - and
tests/test_oracle_vs_reference_golden.py checks this restatement against them.
"""
from __future__ import annotations

from typing import Dict, List, NamedTuple, Optional, Sequence, Tuple

import numpy as np

from . import dsp


class OcclusionOut(NamedTuple):
    importance_map: Optional[np.ndarray]
    baseline_pred: float
    patch_importances: Optional[List[dict]]
    y: np.ndarray
    S: np.ndarray


def occlusion_map(
    y: np.ndarray,
    predictor,
    sr: int,
    n_fft: int = 2048,
    hop_length: int = 512,
    win_length: int = 2048,
    patch_time_frames: int = 1024,
    stride_time_frames: int = 1024,
    patch_freq_percent: float = 5.0,
    stride_freq_percent: float = 5.0,
    occlusion_value: float = 0.0,
    baseline_threshold: float = 0.3,
) -> OcclusionOut:
    y = np.asarray(y, dtype=np.float32)
    S = dsp.stft(y, n_fft, hop_length, win_length).numpy()                      # :603
    baseline_pred = float(predictor.predict(y, sr))                             # :605
    if baseline_pred < baseline_threshold:                                      # :609-619
        return OcclusionOut(None, baseline_pred, None, y, S)
    n_freq, n_time = S.shape
    importance_map = np.zeros((n_freq, n_time))
    count_map = np.zeros((n_freq, n_time))
    t_patch, t_stride = patch_time_frames, stride_time_frames
    patch_freq = max(1, int(round(patch_freq_percent / 100.0 * n_freq)))        # :628-631
    stride_freq = max(1, int(round(stride_freq_percent / 100.0 * n_freq)))
    positions: List[Tuple[int, int]] = []
    for t_start in range(0, max(1, n_time - t_patch + 1), t_stride):            # :644-648
        for f_start in range(0, max(1, n_freq - patch_freq + 1), stride_freq):
            positions.append((t_start, f_start))
    patch_importances: List[dict] = []
    S_occ = S.copy()
    for t_start, f_start in positions:                                          # :665-703
        t_end = min(t_start + t_patch, n_time)
        f_end = min(f_start + patch_freq, n_freq)
        orig = S_occ[f_start:f_end, t_start:t_end].copy()
        S_occ[f_start:f_end, t_start:t_end] = occlusion_value
        y_occ = dsp.istft(S_occ, hop_length, win_length).numpy()
        S_occ[f_start:f_end, t_start:t_end] = orig
        if len(y_occ) > len(y):
            y_occ = y_occ[: len(y)]
        elif len(y_occ) < len(y):
            y_occ = np.pad(y_occ, (0, len(y) - len(y_occ)))
        importance = baseline_pred - float(predictor.predict(y_occ, sr))
        patch_importances.append({"t_start": int(t_start), "t_end": int(t_end), "f_start": int(f_start),
                                  "f_end": int(f_end), "importance": importance})
        importance_map[f_start:f_end, t_start:t_end] += importance
        count_map[f_start:f_end, t_start:t_end] += 1
    importance_map = importance_map / (count_map + 1e-8)                        # :707
    return OcclusionOut(importance_map, baseline_pred, patch_importances, y, S)


def saliency_from_windows(windows: np.ndarray, importances: Sequence[float], n_freq: int, n_time: int) -> np.ndarray:
    """The map/count accumulation of :695-696, :707 alone (float64, window order)."""
    imp_map = np.zeros((n_freq, n_time))
    cnt = np.zeros((n_freq, n_time))
    for (t0, t1, f0, f1), v in zip(np.asarray(windows), importances):
        imp_map[f0:f1, t0:t1] += float(v)
        cnt[f0:f1, t0:t1] += 1
    return imp_map / (cnt + 1e-8)


def _group_metadata(patches: List[dict], top_n: int, file_name: str, group: str, reverse: bool,
                    hop_length: int, sr: int) -> dict:
    srt = sorted(patches, key=lambda p: abs(p["importance"]), reverse=reverse)[:top_n]   # :428-434
    meta = {"file_name": file_name, "group": group, "top_n": int(len(srt)), "windows": []}
    for rank, p in enumerate(srt, 1):
        imp = float(p["importance"])
        meta["windows"].append({
            "rank": int(rank), "t_start": int(p["t_start"]), "t_end": int(p["t_end"]),
            "f_start": int(p["f_start"]), "f_end": int(p["f_end"]),
            "start_time_sec": float(p["t_start"] * hop_length / sr),
            "end_time_sec": float(p["t_end"] * hop_length / sr),
            "importance": imp, "abs_importance": float(abs(imp)),
            "type": "POSITIVE" if imp > 0 else "NEGATIVE" if imp < 0 else "NEUTRAL",
        })
    return meta


def top_window_groups(patches: List[dict], top_n: int, file_name: str, hop_length: int, sr: int) -> Dict[str, dict]:
    """The four JSON payloads of ``_save_top_occlusion_patches_from_list`` (:515-587)."""
    out = {
        "all": _group_metadata(patches, len(patches), file_name, "all", True, hop_length, sr),
        "best": _group_metadata(patches, top_n, file_name, "best", True, hop_length, sr),
        "worst": _group_metadata(patches, top_n, file_name, "worst", False, hop_length, sr),
    }
    pos = sorted([p for p in patches if p["importance"] > 0], key=lambda p: p["importance"], reverse=True)[:top_n]
    neg = sorted([p for p in patches if p["importance"] < 0], key=lambda p: p["importance"], reverse=False)[:top_n]
    mi = pos + neg
    out["most_influential"] = _group_metadata(mi, len(mi), file_name, "most_influential", False, hop_length, sr)
    return out


def window_audio(y: np.ndarray, S: np.ndarray, p: dict, hop_length: int, win_length: int,
                 use_original_audio: bool) -> np.ndarray:
    """Audio written for one top window (:456-483)."""
    t0, t1, f0, f1 = p["t_start"], p["t_end"], p["f_start"], p["f_end"]
    window_samples = max(1, (t1 - t0) * hop_length)
    start = int(t0 * hop_length)
    if use_original_audio:
        w = y[start: min(start + window_samples, len(y))]
        if len(w) < window_samples:
            w = np.pad(w, (0, window_samples - len(w)))
        return w
    masked = np.zeros_like(S)
    masked[f0:f1, t0:t1] = S[f0:f1, t0:t1]
    full = dsp.istft(masked, hop_length, win_length).numpy()
    return full[start: min(start + window_samples, len(full))]


class FBPOut(NamedTuple):
    importance_map: np.ndarray
    baseline_pred: float
    batch_importances: List[dict]
    S: np.ndarray


def fbp_component(
    sig: np.ndarray,
    predictor,
    sr: int,
    bands: Sequence[Tuple[float, float]],
    attenuation: float,
    transition_mode: str = "rel",
    transition_hz: float = 0.0,
    transition_rel: float = 0.0,
    transition_min_hz: float = 0.0,
    transition_max_hz: float = 0.0,
    n_fft: int = 2048,
    hop_length: int = 512,
    win_length: int = 2048,
    normalize_loudness: bool = True,
    component_name: str = "mixture",
    return_audio: bool = False,
):
    sig = np.asarray(sig)
    orig_prob = float(predictor.predict(sig, sr))                               # :544
    S = dsp.stft(sig, n_fft, hop_length, win_length).numpy()                    # :565
    mag, phase = dsp.magphase(S)                                                # :566
    freqs = dsp.fft_frequencies(sr, n_fft)                                      # :567
    batch: List[dict] = []
    importance_map = np.zeros_like(mag, dtype=float)
    audio = []
    for (low, high) in bands:                                                   # :573-653
        bw = float(high - low)
        if transition_mode == "rel":
            trans = float(np.clip(bw * transition_rel, transition_min_hz, transition_max_hz))
        else:
            trans = float(transition_hz)
        keep = _keep_mask(freqs, low, high, trans)
        keep_band = keep + attenuation * (1.0 - keep)
        S_p = (mag * keep_band[:, None]) * phase                                # float64 x complex64 -> complex128
        y_p = dsp.istft(S_p, hop_length, win_length).numpy()
        if normalize_loudness:
            y_p = dsp.match_rms(sig, y_p)
        if return_audio:
            audio.append(y_p)
        delta = float(orig_prob - float(predictor.predict(y_p, sr)))
        batch.append({"component": component_name, "low": float(low), "high": float(high), "importance": delta})
        importance_map[(freqs >= low) & (freqs <= high), :] += delta
    out = FBPOut(importance_map, orig_prob, batch, S)
    return (out, audio) if return_audio else out


def _keep_mask(freqs, low, high, trans):
    """src/dsp_band_ops.py:236-259 restated."""
    f = freqs.astype(float)
    m = np.ones_like(f, dtype=float)
    m[(f >= low) & (f <= high)] = 0.0
    if trans > 0:
        tl = (f >= (low - trans)) & (f < low)
        if np.any(tl):
            m[tl] = 0.5 * (1.0 + np.cos(np.pi * ((f[tl] - (low - trans)) / trans)))
        th = (f > high) & (f <= (high + trans))
        if np.any(th):
            m[th] = 0.5 * (1.0 + np.cos(np.pi * (1.0 - (f[th] - high) / trans)))
    return np.clip(m, 0.0, 1.0)


def stem_mask_probs(stems: np.ndarray, masks: np.ndarray, predictor, sr: int) -> np.ndarray:
    """``predict_fn_unified`` over LIME stem recombinations: x = sum_i m_i * stem_i -> [1-p, p]
    (src/lime_explainer.py:283-301; composition per audioLIME's SpleeterFactorization.compose_model_input)."""
    out = np.zeros((masks.shape[0], 2))
    for i, m in enumerate(masks):
        x = np.zeros(stems.shape[1], dtype=np.float32)
        for j, on in enumerate(m):
            if on:
                x = x + stems[j]
        p = float(predictor.predict(x, sr))
        out[i] = (1.0 - p, p)
    return out


# ------------------------------------------------------------------------------------------------- RISE (:722-806)
def _lowbias32(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint64)
    x ^= x >> np.uint64(16); x = (x * np.uint64(0x7FEB352D)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(15); x = (x * np.uint64(0x846CA68B)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(16)
    return x


def rise_keep_mask(seed: int, mask_index: int, n_freq: int, n_time: int, keep_probability: float) -> np.ndarray:
    """Keep mask ``[n_freq, n_time]`` (bool) of one RISE iteration.  The reference's ``np.random.rand(n_freq, n_time) >
    1 - p`` (:768) uses the unseeded global RNG; the build defines the bits as a counter-based hash of (seed, mask, cell)
    (csrc/common.h: rise_mask_key / rise_keep) - restated here in numpy so that the CPU loop sees the very same masks."""
    m32 = np.uint64(0xFFFFFFFF)
    key = _lowbias32(np.array([(np.uint64(seed & 0xFFFFFFFF) * np.uint64(0x9E3779B9) + np.uint64(mask_index) * np.uint64(0x85EBCA6B)
                                + np.uint64(0x165667B1)) & m32]))[0]
    t = np.arange(n_time, dtype=np.uint64)[None, :]
    f = np.arange(n_freq, dtype=np.uint64)[:, None]
    cell = (t * np.uint64(1025) + f) & m32
    u = _lowbias32(key ^ ((cell * np.uint64(0xC2B2AE35)) & m32))
    v = keep_probability * 4294967296.0
    thr = 4294967295 if v >= 4294967295.0 else (0 if v <= 0 else int(v))
    return u < np.uint64(thr)


class RiseOut(NamedTuple):
    importance_map: Optional[np.ndarray]
    raw_map: Optional[np.ndarray]
    baseline_pred: float
    predictions: Optional[List[float]]


def rise_map(y: np.ndarray, predictor, sr: int, n_masks: int, mask_probability: float, seed: int, n_fft: int = 2048,
             hop_length: int = 512, win_length: int = 2048, baseline_threshold: float = 0.3) -> RiseOut:
    """``_compute_rise_map`` (:722-806) with the build's mask bits in place of ``np.random.rand``."""
    y = np.asarray(y, dtype=np.float32)
    S = dsp.stft(y, n_fft, hop_length, win_length).numpy()
    baseline_pred = float(predictor.predict(y, sr))
    if baseline_pred < baseline_threshold:                                      # :741-744
        return RiseOut(None, None, baseline_pred, None)
    n_freq, n_time = S.shape
    importance_map = np.zeros((n_freq, n_time))
    predictions: List[float] = []
    for mask_idx in range(n_masks):                                             # :766-790
        mask = rise_keep_mask(seed, mask_idx, n_freq, n_time, mask_probability).astype(float)
        y_masked = dsp.istft(S * mask, hop_length, win_length).numpy()
        if len(y_masked) > len(y):
            y_masked = y_masked[: len(y)]
        elif len(y_masked) < len(y):
            y_masked = np.pad(y_masked, (0, len(y) - len(y_masked)))
        masked_pred = float(predictor.predict(y_masked, sr))
        predictions.append(masked_pred)
        importance_map += mask * masked_pred
    raw = importance_map / (n_masks * mask_probability + 1e-8)                  # :798
    norm = (raw - raw.min()) / (raw.max() - raw.min() + 1e-8)                   # :801
    return RiseOut(norm, raw, baseline_pred, predictions)
