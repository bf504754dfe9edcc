"""Host-side integer / indexing logic of the perturbation hot path.

Everything here is bit-exact integer or float64 arithmetic that the reference performs in
Python/numpy on the host; it stays on the host in this engine as well (it is O(N) for N windows and
must reproduce Python's ``round`` (IEEE double, round-half-to-even) exactly).

Reference citations (relative to /root/reference):
  * patch grid ...................... src/spectrogram_explainability.py:621-648, 667-668
  * FREQUENCY_BAND_PRESETS .......... src/dsp_band_ops.py:212-226
  * smooth_band_keep_mask ........... src/dsp_band_ops.py:236-259
  * _band_transition_width .......... src/dsp_band_ops.py:428-435
  * keep_band / band->bin rows ...... src/dsp_band_ops.py:576, 652-653
  * top-k window groups ............. src/spectrogram_explainability.py:413-587
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np

# Same preset table as the reference (values are the contract, src/dsp_band_ops.py:212-226).
FREQUENCY_BAND_PRESETS: Dict[str, List[Tuple[int, int]]] = {
    "default": [(20, 100), (100, 250), (250, 2000), (2000, 4000), (4000, 8000), (8000, 16000)],
    "detailed_voice": [
        (20, 60), (60, 250), (250, 500), (500, 2000), (2000, 4000), (4000, 6000), (6000, 12000), (12000, 21000),
    ],
    "high_resolution": [
        (20, 60), (60, 100), (100, 250), (250, 500), (500, 1000), (1000, 2000), (2000, 4000), (4000, 6000),
        (6000, 8000), (8000, 10000), (10000, 12000), (12000, 16000), (16000, 21000),
    ],
}


def stft_shape(n_samples: int, n_fft: int, hop_length: int) -> Tuple[int, int]:
    """(n_freq, n_time) of a centred STFT (librosa semantics: ``1 + n_fft//2``, ``1 + L//hop``)."""
    return 1 + n_fft // 2, 1 + n_samples // hop_length


def occlusion_patch_sizes(n_freq: int, patch_freq_percent: float, stride_freq_percent: float) -> Tuple[int, int]:
    """``max(1, int(round(pct/100*n_freq)))`` with Python's banker's rounding (:628-631)."""
    patch_freq = max(1, int(round(patch_freq_percent / 100.0 * n_freq)))
    stride_freq = max(1, int(round(stride_freq_percent / 100.0 * n_freq)))
    return patch_freq, stride_freq


def occlusion_windows(
    n_freq: int,
    n_time: int,
    patch_time_frames: int,
    stride_time_frames: int,
    patch_freq_percent: float,
    stride_freq_percent: float,
) -> np.ndarray:
    """Window list ``int32[N, 4] = (t_start, t_end, f_start, f_end)`` in the reference's order
    (t-major, f-minor; ends clipped to the spectrogram, :644-648 and :667-668)."""
    patch_freq, stride_freq = occlusion_patch_sizes(n_freq, patch_freq_percent, stride_freq_percent)
    t_patch, t_stride = int(patch_time_frames), int(stride_time_frames)
    rows = []
    for t_start in range(0, max(1, n_time - t_patch + 1), t_stride):
        for f_start in range(0, max(1, n_freq - patch_freq + 1), stride_freq):
            rows.append((t_start, min(t_start + t_patch, n_time), f_start, min(f_start + patch_freq, n_freq)))
    return np.asarray(rows, dtype=np.int32).reshape(-1, 4)


def fft_frequencies(sr: float, n_fft: int) -> np.ndarray:
    """``librosa.fft_frequencies`` == ``np.fft.rfftfreq(n_fft, 1/sr)`` (float64)."""
    return np.fft.rfftfreq(n=n_fft, d=1.0 / sr)


def band_transition_width(
    low: float, high: float, mode: str, rel: float, min_hz: float, max_hz: float, hz: float
) -> float:
    bw = float(high - low)
    if mode == "rel":
        return float(np.clip(bw * rel, min_hz, max_hz))
    return float(hz)


def smooth_band_keep_mask(freqs: np.ndarray, low: float, high: float, trans: float = 200.0) -> np.ndarray:
    """keep = 0 inside [low, high], raised-cosine ramps of width ``trans`` either side, 1 elsewhere."""
    f = np.asarray(freqs, dtype=float)
    m = np.ones_like(f)
    m[(f >= low) & (f <= high)] = 0.0
    if trans > 0:
        tl = (f >= (low - trans)) & (f < low)
        m[tl] = 0.5 * (1.0 + np.cos(np.pi * ((f[tl] - (low - trans)) / trans)))
        th = (f > high) & (f <= (high + trans))
        m[th] = 0.5 * (1.0 + np.cos(np.pi * (1.0 - (f[th] - high) / trans)))
    return np.clip(m, 0.0, 1.0)


def band_gain_table(
    bands: Sequence[Tuple[float, float]],
    sr: float,
    n_fft: int,
    attenuation: float,
    transition_mode: str = "rel",
    transition_rel: float = 0.0,
    transition_min_hz: float = 0.0,
    transition_max_hz: float = 0.0,
    transition_hz: float = 0.0,
) -> np.ndarray:
    """Per-band gain over STFT bins, float64 ``[n_bands, n_freq]``:
    ``keep + attenuation * (1 - keep)`` (src/dsp_band_ops.py:574-576)."""
    freqs = fft_frequencies(sr, n_fft)
    out = np.empty((len(bands), freqs.shape[0]), dtype=np.float64)
    for i, (low, high) in enumerate(bands):
        trans = band_transition_width(low, high, transition_mode, transition_rel, transition_min_hz,
                                      transition_max_hz, transition_hz)
        keep = smooth_band_keep_mask(freqs, low, high, trans=trans)
        out[i] = keep + attenuation * (1.0 - keep)
    return out


def band_bin_ranges(bands: Sequence[Tuple[float, float]], sr: float, n_fft: int) -> np.ndarray:
    """Hard inclusive band -> bin rows, ``int32[n_bands, 2] = (first_bin, last_bin_exclusive)``;
    an empty band gives ``(0, 0)`` (mask ``(freqs >= low) & (freqs <= high)``, :652)."""
    freqs = fft_frequencies(sr, n_fft)
    out = np.zeros((len(bands), 2), dtype=np.int32)
    for i, (low, high) in enumerate(bands):
        idx = np.nonzero((freqs >= low) & (freqs <= high))[0]
        if idx.size:
            out[i] = (idx[0], idx[-1] + 1)  # the mask is contiguous because freqs is monotone
    return out


def band_map_view(band_rows: np.ndarray, deltas: Sequence[float], n_freq: int, n_time: int) -> np.ndarray:
    """The FBP importance map ``map[(freqs >= low) & (freqs <= high), :] += delta`` (src/dsp_band_ops.py:652-653) as a READ-ONLY
    broadcast view: every column of that map is the same ``[n_freq]`` float64 vector (the additions run in band order, like
    the reference's), so the ``[n_freq, n_time]`` array (30.8 MB per 120 s track) never has to be produced or copied.
    ``np.array(view)`` materialises a writable copy."""
    col = np.zeros(n_freq, dtype=np.float64)
    for (r0, r1), d in zip(np.asarray(band_rows).reshape(-1, 2), deltas):
        col[int(r0):int(r1)] += np.float64(d)
    return np.broadcast_to(col[:, None], (n_freq, n_time))


def importance_type(v: float) -> str:
    return "POSITIVE" if v > 0 else "NEGATIVE" if v < 0 else "NEUTRAL"


def stable_order(keys: np.ndarray, descending: bool) -> np.ndarray:
    """Index order of Python's ``sorted(..., key=, reverse=descending)``: stable, and ``reverse=True``
    also keeps equal keys in their original order."""
    keys = np.asarray(keys, dtype=np.float64)
    if descending:
        return np.argsort(-keys, kind="stable")
    return np.argsort(keys, kind="stable")


def snap_ties(importances: Sequence[float], tie_epsilon: float) -> np.ndarray:
    """Ranking keys with ``|v| < tie_epsilon`` snapped to exactly 0.0 (builder extension, off at 0.0).

    The reference ranks the raw floats (:428-434, 566-571).  Windows whose occlusion changes nothing (e.g. bins above a
    track's band limit) have ``|delta|`` at the arithmetic noise floor of the classifier (exactly 0.0 in fp32 eager
    PyTorch, ~1e-5 with bf16 GEMM inputs), so their mutual order - which decides the *worst* group - is not reproducible
    between two arithmetics, including reference CPU vs reference CUDA.  With a ``tie_epsilon`` above that noise floor such
    windows are exact ties and keep grid order (Python's stable sort), on every implementation."""
    imp = np.array(importances, dtype=np.float64)
    if tie_epsilon > 0.0:
        imp[np.abs(imp) < tie_epsilon] = 0.0
    return imp


def topk_boundary_margins(importances: Sequence[float], top_n: int) -> Dict[str, float]:
    """Gap between the last key inside each top-``top_n`` group and the first key outside it (inf if nothing is outside).
    A group is reproducible across arithmetics whose per-window error is below half of its margin."""
    imp = np.asarray(importances, dtype=np.float64)
    a_desc = np.sort(np.abs(imp))[::-1]
    a_asc = a_desc[::-1]
    pos = np.sort(imp[imp > 0])[::-1]
    neg = np.sort(imp[imp < 0])

    def gap(keys, asc):
        if len(keys) <= top_n or top_n <= 0:
            return float("inf")
        return float(keys[top_n] - keys[top_n - 1]) if asc else float(keys[top_n - 1] - keys[top_n])

    return {"best": gap(a_desc, False), "worst": gap(a_asc, True),
            "most_influential": min(gap(pos, False), gap(neg, True))}


def topk_window_groups(importances: Sequence[float], top_n: int, tie_epsilon: float = 0.0) -> Dict[str, np.ndarray]:
    """Indices (into the window list) of the reference's four groups, in output (rank) order.

    all: abs desc; best: top_n abs desc; worst: top_n abs asc; most_influential: top_n positives by
    value desc ++ top_n negatives by value asc, the concatenation re-sorted by abs asc (:515-587).
    ``tie_epsilon`` > 0 ranks ``snap_ties(importances, tie_epsilon)`` instead of the raw values.
    """
    imp = snap_ties(importances, tie_epsilon)
    a = np.abs(imp)
    all_desc = stable_order(a, True)
    asc = stable_order(a, False)
    pos = np.nonzero(imp > 0)[0]
    neg = np.nonzero(imp < 0)[0]
    top_pos = pos[stable_order(imp[pos], True)][:top_n]
    top_neg = neg[stable_order(imp[neg], False)][:top_n]
    mi = np.concatenate([top_pos, top_neg]).astype(np.int64)
    mi = mi[stable_order(a[mi], False)] if mi.size else mi
    return {
        "all": all_desc.astype(np.int64),
        "best": all_desc[:top_n].astype(np.int64),
        "worst": asc[:top_n].astype(np.int64),
        "most_influential": mi,
    }


def shard_range(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous balanced slice of ``range(n_items)`` owned by ``rank`` (first ``n % world`` ranks get
    one extra item)."""
    base, rem = divmod(n_items, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)
