"""Occlusion explainer - drop-in for ``SpectrogramExplainability`` (src/spectrogram_explainability.py:288-1049).

Same constructor keywords, same ``_compute_occlusion_map`` / ``process_audio_file`` / ``run_experiment`` entry
points, same ``OcclusionResult`` and on-disk layout (``top_windows/<group>/<file>__<group>_occlusion_patches_from_list
.json`` + WAVs).  The reference's hot loop (:665-703: zero a patch, iSTFT, predict, accumulate - one copy at a time)
is replaced by ONE batched device sweep over all windows, sharded across ranks when torch.distributed is up.
"""
from __future__ import annotations

import json
import os
from datetime import datetime
from pathlib import Path
from typing import Any, Dict, NamedTuple, Optional

import numpy as np

from . import dist, grid, mel_host
from .audio_io import AudioDecodeError, load_audio, write_wav
from .sonics_api import B200Predictor


class OcclusionResult(NamedTuple):
    importance_map: Optional[np.ndarray]
    spectrogram_db: np.ndarray
    baseline_pred: float
    y: np.ndarray
    S: np.ndarray
    patch_importances: Optional[list]


class RiseResult(NamedTuple):
    importance_map: Optional[np.ndarray]
    spectrogram_db: np.ndarray
    baseline_pred: float
    y: np.ndarray
    S: np.ndarray


def amplitude_to_db_refmax(S: np.ndarray) -> np.ndarray:
    """``librosa.amplitude_to_db(np.abs(S), ref=np.max)`` (visualisation only, :387)."""
    mag = np.abs(S).astype(np.float32)
    ref = float(mag.max()) ** 2
    log_spec = 10.0 * np.log10(np.maximum(1e-10, mag ** 2)) - 10.0 * np.log10(max(1e-10, ref))
    return np.maximum(log_spec, log_spec.max() - 80.0)


class SpectrogramCheckpoint:
    """Per-file resume bookkeeping, same JSON as the reference (:97-135)."""

    def __init__(self, checkpoint_dir):
        self.checkpoint_dir = Path(checkpoint_dir)
        self.checkpoint_dir.mkdir(parents=True, exist_ok=True)
        self.checkpoint_file = self.checkpoint_dir / "spectrogram_checkpoint.json"
        self.progress_log = self.checkpoint_dir / "spectrogram_progress.txt"

    def load_processed_files(self) -> set:
        if self.checkpoint_file.exists():
            with open(self.checkpoint_file, "r", encoding="utf-8") as f:
                return set(json.load(f).get("processed_files", []))
        return set()

    def mark_as_processed(self, file_path: str) -> None:
        data = {"processed_files": [], "last_updated": None}
        if self.checkpoint_file.exists():
            with open(self.checkpoint_file, "r", encoding="utf-8") as f:
                data = json.load(f)
        if file_path not in data["processed_files"]:
            data["processed_files"].append(file_path)
        data["last_updated"] = datetime.now().isoformat()
        with open(self.checkpoint_file, "w", encoding="utf-8") as f:
            json.dump(data, f, ensure_ascii=False, indent=2)
        with open(self.progress_log, "a", encoding="utf-8") as f:
            f.write(f"[PROCESSED] {datetime.now().isoformat()} | {file_path}\n")


def append_update_spectrogram_results(new_results: dict, results_path: Path) -> None:
    """Merge ``{folder: {file: result}}`` into the results JSON (:253-286)."""
    merged: dict = {}
    if results_path.exists():
        try:
            with open(results_path, "r", encoding="utf-8") as f:
                merged = json.load(f)
        except Exception:
            merged = {}
    for folder, files in new_results.items():
        merged.setdefault(folder, {}).update(files)
    results_path.parent.mkdir(parents=True, exist_ok=True)
    with open(results_path, "w", encoding="utf-8") as f:
        json.dump(merged, f, indent=4, ensure_ascii=False)


class SpectrogramExplainability:
    def __init__(self, predictor, sr: int = 44100, duration: int = 120, n_fft: int = 2048, hop_length: int = 512,
                 win_length: int = 2048, n_mels: int = 128, n_iter: int = 256, spec_type: str = "mel",
                 fmax: Optional[float] = None, top_n_windows: int = 5, method: str = "rise",
                 use_original_audio: bool = True, patch_time_frames: int = 2048, stride_time_frames: int = 2048,
                 patch_freq_percent: float = 25.0, stride_freq_percent: float = 25.0, n_masks: int = 500,
                 mask_probability: float = 0.5, checkpoint_dir=None, highlight_percent: float = 20.0,
                 abs_threshold: float = 0.0, rise_seed: int = 0, tie_epsilon: float = 0.0, mel_seed: int = 0,
                 nnls_iter: int = 16):
        if not isinstance(predictor, B200Predictor):
            raise TypeError("the B200 occlusion sweep needs a B200Predictor (the classifier runs inside the sweep); "
                            f"got {type(predictor).__name__}")
        self.predictor = predictor
        self.sr, self.duration = sr, duration
        self.n_fft, self.hop_length, self.win_length = n_fft, hop_length, win_length
        self.n_mels, self.n_iter = n_mels, n_iter
        self.top_n_windows = top_n_windows
        self.method = method.lower()
        self.spec_type = spec_type.lower()
        if self.spec_type not in ["mel", "stft"]:
            raise ValueError(f"Unsupported spec_type: {spec_type}. Use 'mel' or 'stft'.")
        self.fmax = fmax if fmax is not None else sr // 2
        self.patch_time_frames, self.stride_time_frames = patch_time_frames, stride_time_frames
        self.patch_freq_percent, self.stride_freq_percent = patch_freq_percent, stride_freq_percent
        self.use_original_audio = use_original_audio
        self.n_masks, self.mask_probability = n_masks, mask_probability
        self.speculative_baseline = False   # True: always evaluate the baseline inside the sweep (see occlusion_map_from_wave)
        self.rise_seed = rise_seed          # the reference draws RISE masks from the unseeded numpy global RNG (:768)
        self.tie_epsilon = float(tie_epsilon)   # 0.0 = rank the raw importances like the reference; see grid.snap_ties
        # spec_type='mel': the reference inverts with librosa's L-BFGS-B NNLS + Griffin-Lim from UNSEEDED random phases
        # (:394-402), which no two runs reproduce; the build's inverse is seeded and its NNLS is the projected-gradient
        # iteration of csrc/mel_domain.cu (oracle/mel.py restates both)
        self.mel_seed, self.nnls_iter = int(mel_seed), int(nnls_iter)
        self._mel_key = None
        self.highlight_percent, self.abs_threshold = highlight_percent, abs_threshold
        self.checkpoint = SpectrogramCheckpoint(checkpoint_dir) if checkpoint_dir else None

    @classmethod
    def from_config(cls, config, predictor, checkpoint_dir=None) -> "SpectrogramExplainability":
        """Explainer from a reference YAML file / dict, keys and fallbacks as run_spectrogram_experiment.py:157-205."""
        from .config import spectrogram_explainer_kwargs
        return cls(predictor=predictor, **spectrogram_explainer_kwargs(config, checkpoint_dir))

    # -- guards for the variants that have no reference parity ----------------
    def _require_stft_occlusion(self, method: str = "occlusion", allow_mel: bool = False) -> None:
        if self.spec_type != "stft" and not allow_mel:
            raise NotImplementedError("spec_type='mel' is implemented for the occlusion method only")
        if self.method != method:
            raise NotImplementedError(f"this explainer was built with method={self.method!r}; "
                                      f"{'_compute_rise_map' if self.method == 'rise' else '_compute_occlusion_map'} is its entry point")
        if (self.n_fft, self.hop_length, self.win_length) != (2048, 512, 2048):
            raise NotImplementedError("the CUDA STFT/iSTFT kernels are built for n_fft=2048, hop=512, win=2048")

    def _predict_fn(self, waveform: np.ndarray, sr: int) -> float:
        return float(self.predictor.predict(waveform, sr))          # errors propagate; no 0.0 fallback

    def _compute_spectrogram(self, y: np.ndarray):
        """(S, S_db) on the GPU: complex64 STFT ``[1025, 1 + len(y)//512]`` (librosa.stft semantics) or, for
        ``spec_type='mel'``, the float32 power mel spectrogram ``[n_mels, n_time]`` (librosa.feature.melspectrogram, :367-377)."""
        self._require_stft_occlusion(self.method, allow_mel=True)
        eng = self.predictor.engine
        eng.set_track(y)
        if self.spec_type == "mel":
            self._ensure_mel_basis()
            S = eng.mel_spectrogram()
            return S, mel_host.power_to_db_refmax(S)
        S = eng.spectrogram()
        return S, amplitude_to_db_refmax(S)

    def _ensure_mel_basis(self) -> None:
        """Upload the Slaney filterbank + NNLS operators once per configuration.  Like the reference, the forward filterbank
        honours ``fmax`` (:375) while the inverse call passes none (:395-402): both are the same bank unless fmax < sr/2, in
        which case - as in the reference - the inversion uses the [0, sr/2] bank."""
        key = (self.sr, self.n_fft, self.n_mels, float(self.fmax))
        if self._mel_key == key:
            return
        if float(self.fmax) != float(self.sr // 2) and float(self.fmax) != self.sr / 2:
            raise NotImplementedError("spec_type='mel' with fmax != sr/2 pairs two different filterbanks in the reference "
                                      "(forward :375 vs inverse :395-402); not supported")
        basis = mel_host.mel_filterbank(self.sr, self.n_fft, self.n_mels, 0.0, None)
        pinv, step = mel_host.nnls_operators(basis)
        self.predictor.engine.set_mel_basis(basis, pinv, step)
        self._mel_key = key

    # -- the hot path ---------------------------------------------------------------------------------
    def occlusion_map_from_wave(self, y: np.ndarray, occlusion_value: float = 0.0, baseline_threshold: float = 0.3,
                                verbose: bool = True, want_spectrogram: bool = True) -> OcclusionResult:
        self._require_stft_occlusion(allow_mel=True)
        if self.spec_type == "mel":
            return self._mel_occlusion_map_from_wave(y, occlusion_value, baseline_threshold, verbose)
        eng = self.predictor.engine
        y = np.ascontiguousarray(np.asarray(y, dtype=np.float32))
        eng.set_track(y)
        S = eng.spectrogram() if want_spectrogram else None
        S_db = amplitude_to_db_refmax(S) if want_spectrogram else None
        n_freq, n_time = eng.track_shape()
        windows = grid.occlusion_windows(n_freq, n_time, self.patch_time_frames, self.stride_time_frames,
                                         self.patch_freq_percent, self.stride_freq_percent)
        _, world = dist.world()
        # The reference predicts the baseline first and skips the file below the threshold (:605-619).  When nothing can be
        # skipped (threshold <= 0) - or the caller accepts a wasted sweep on skipped files (speculative_baseline) - the
        # track rides in the sweep's own device pass as one more copy; the value is predict_track()'s, bit for bit.
        fused = world == 1 and (baseline_threshold <= 0.0 or self.speculative_baseline)
        if fused:
            probs, base = eng.occlusion_sweep(windows, occlusion_value, with_baseline=True)
            baseline_pred = float(base)
        else:
            baseline_pred = float(eng.predict_track())
        if verbose:
            print(f"    Baseline prediction: {baseline_pred:.4f}")
        if baseline_pred < baseline_threshold:
            return OcclusionResult(None, S_db, baseline_pred, y, S, None)
        if verbose:
            print(f"    Processing {len(windows)} patches (t_patch={self.patch_time_frames}, "
                  f"t_stride={self.stride_time_frames}) in one batched sweep...")
        if not fused:
            probs = dist.sharded_sweep(lambda w: eng.occlusion_sweep(w, occlusion_value), windows)
        # importance = baseline_pred - occluded_pred on Python floats (:684)
        importances = [baseline_pred - float(p) for p in probs]
        patch_importances = [
            {"t_start": int(w[0]), "t_end": int(w[1]), "f_start": int(w[2]), "f_end": int(w[3]), "importance": imp}
            for w, imp in zip(windows, importances)
        ]
        importance_map = eng.saliency_map(windows, np.asarray(importances, dtype=np.float64))
        if verbose:
            print(f"    Completed | Mean importance: {importance_map.mean():.4f}, Max: {importance_map.max():.4f}")
        return OcclusionResult(importance_map, S_db, baseline_pred, y, S, patch_importances)

    def _mel_occlusion_map_from_wave(self, y, occlusion_value, baseline_threshold, verbose) -> OcclusionResult:
        """The occlusion loop over the MEL spectrogram (:663-703 with spec_type == 'mel'): window i zeroes mel bins
        ``[f0, f1)`` of frames ``[t0, t1)``, the result is inverted (NNLS -> Griffin-Lim, ``n_iter`` iterations, phases of
        index i) and classified.  One batched device sweep; the windows are sharded across ranks like the STFT ones."""
        eng = self.predictor.engine
        y = np.ascontiguousarray(np.asarray(y, dtype=np.float32))
        S, S_db = self._compute_spectrogram(y)
        n_freq, n_time = S.shape
        windows = grid.occlusion_windows(n_freq, n_time, self.patch_time_frames, self.stride_time_frames,
                                         self.patch_freq_percent, self.stride_freq_percent)
        baseline_pred = float(eng.predict_track())
        if verbose:
            print(f"    Baseline prediction: {baseline_pred:.4f}")
        if baseline_pred < baseline_threshold:
            return OcclusionResult(None, S_db, baseline_pred, y, S, None)
        rank, world = dist.world()
        lo, hi = grid.shard_range(len(windows), rank, world)
        local = (eng.mel_sweep(eng.MASK_OCCLUDE, windows[lo:hi], self.n_iter, self.nnls_iter, self.mel_seed, first_index=lo,
                               occlusion_value=occlusion_value) if hi > lo else np.zeros(0, np.float32))
        probs = dist.gather_shards(local, len(windows))
        importances = [baseline_pred - float(p) for p in probs]
        patch_importances = [
            {"t_start": int(w[0]), "t_end": int(w[1]), "f_start": int(w[2]), "f_end": int(w[3]), "importance": imp}
            for w, imp in zip(windows, importances)
        ]
        importance_map = eng.saliency_map_shape(windows, np.asarray(importances, dtype=np.float64), n_freq, n_time)
        return OcclusionResult(importance_map, S_db, baseline_pred, y, S, patch_importances)

    # -- RISE (:722-806) -------------------------------------------------------------------------------
    def rise_map_from_wave(self, y: np.ndarray, baseline_threshold: float = 0.3, verbose: bool = True) -> RiseResult:
        """``n_masks`` i.i.d. Bernoulli(``mask_probability``) keep masks over the STFT, each -> iSTFT -> prediction;
        ``map = sum_i mask_i * pred_i / (n_masks * p + 1e-8)`` scaled to [0, 1].  The masks never exist in memory: the iSTFT
        load stage and the map reduction both evaluate the same counter-based hash of (``rise_seed``, mask, cell).  The
        reference draws them from numpy's unseeded global RNG, so individual runs are not comparable bit for bit even
        reference-vs-reference; parity is against ``oracle/loops.py::rise_map`` with the same hash."""
        self._require_stft_occlusion("rise")
        eng = self.predictor.engine
        y = np.ascontiguousarray(np.asarray(y, dtype=np.float32))
        eng.set_track(y)
        S = eng.spectrogram()
        S_db = amplitude_to_db_refmax(S)
        baseline_pred = float(eng.predict_track())
        if verbose:
            print(f"    Baseline prediction: {baseline_pred:.4f}")
        if baseline_pred < baseline_threshold:
            return RiseResult(None, S_db, baseline_pred, y, S)
        rank, world = dist.world()
        lo, hi = grid.shard_range(self.n_masks, rank, world)            # masks are independent: shard them like windows
        local = eng.rise_sweep(hi - lo, self.rise_seed, self.mask_probability, first_mask=lo) if hi > lo else np.zeros(0, np.float32)
        preds = dist.gather_shards(local, self.n_masks)
        importance_map = eng.rise_map(np.array([float(p) for p in preds], dtype=np.float64), self.rise_seed, self.mask_probability)
        importance_map = (importance_map - importance_map.min()) / (importance_map.max() - importance_map.min() + 1e-8)
        if verbose:
            print(f"    Completed | Mean importance: {importance_map.mean():.4f}, Max: {importance_map.max():.4f}")
        return RiseResult(importance_map, S_db, baseline_pred, y, S)

    def _compute_rise_map(self, audio_path: str, baseline_threshold: float = 0.3, verbose: bool = True) -> RiseResult:
        y, _ = load_audio(audio_path, sr=self.sr, duration=self.duration, mono=True, resample=self.predictor.engine.resample)
        return self.rise_map_from_wave(y, baseline_threshold, verbose)

    def _compute_occlusion_map(self, audio_path: str, occlusion_value: float = 0.0, baseline_threshold: float = 0.3,
                               verbose: bool = True) -> OcclusionResult:
        y, _ = load_audio(audio_path, sr=self.sr, duration=self.duration, mono=True, resample=self.predictor.engine.resample)
        return self.occlusion_map_from_wave(y, occlusion_value, baseline_threshold, verbose)

    # -- top-k windows -> JSON / WAV ----------------------------------------------------------------------
    def top_window_groups(self, patch_importances: list, top_n: int, file_name: str) -> Dict[str, dict]:
        """Metadata payloads of the four groups (all / best / worst / most_influential), ranks as in :413-587."""
        raw = np.array([p["importance"] for p in patch_importances], dtype=np.float64)
        imp = grid.snap_ties(raw, self.tie_epsilon)      # ranking keys only; the JSON reports the raw values
        eng = self.predictor.engine
        order_desc = eng.rank(imp, 0)                    # |imp| descending, stable
        order_asc = eng.rank(imp, 1)                     # |imp| ascending, stable
        by_val_desc = eng.rank(imp, 2)
        by_val_asc = eng.rank(imp, 3)
        top_pos = [i for i in by_val_desc if imp[i] > 0][:top_n]
        top_neg = [i for i in by_val_asc if imp[i] < 0][:top_n]
        mi = top_pos + top_neg
        mi = [mi[j] for j in np.argsort(np.abs(imp[mi]), kind="stable")] if mi else []
        groups = {"all": list(order_desc), "best": list(order_desc[:top_n]), "worst": list(order_asc[:top_n]),
                  "most_influential": mi}
        out = {}
        for g, idx in groups.items():
            meta = {"file_name": file_name, "group": g, "top_n": int(len(idx)), "windows": []}
            for rank, i in enumerate(idx, 1):
                p = patch_importances[int(i)]
                v = float(p["importance"])
                meta["windows"].append({
                    "rank": int(rank), "t_start": int(p["t_start"]), "t_end": int(p["t_end"]),
                    "f_start": int(p["f_start"]), "f_end": int(p["f_end"]),
                    "start_time_sec": float(p["t_start"] * self.hop_length / self.sr),
                    "end_time_sec": float(p["t_end"] * self.hop_length / self.sr),
                    "importance": v, "abs_importance": float(abs(v)), "type": grid.importance_type(v),
                })
            out[g] = meta
        return out

    def window_audio(self, y: np.ndarray, windows_meta: list) -> list:
        """Audio of each listed window: a slice of ``y`` (use_original_audio) or the patch-only iSTFT (:456-483)."""
        out = []
        if not windows_meta:
            return out
        if not self.use_original_audio:
            w = np.array([[m["t_start"], m["t_end"], m["f_start"], m["f_end"]] for m in windows_meta], np.int32)
            if self.spec_type == "mel":                  # masked_S = zeros except the patch -> mel_to_audio -> slice (:472-483)
                eng = self.predictor.engine
                audio = eng.mel_sweep(eng.MASK_KEEP_ONLY, w, self.n_iter, self.nnls_iter, self.mel_seed, first_index=1 << 20,
                                      want_prob=False, want_audio=True)
                full = []
                for m, a in zip(windows_meta, audio):
                    start = int(m["t_start"] * self.hop_length)
                    n = max(1, (m["t_end"] - m["t_start"]) * self.hop_length)
                    full.append(a[start: min(start + n, len(a))])
            else:
                full = self.predictor.engine.window_audio(w)
        for k, m in enumerate(windows_meta):
            n = max(1, (m["t_end"] - m["t_start"]) * self.hop_length)
            start = int(m["t_start"] * self.hop_length)
            if self.use_original_audio:
                seg = y[start: min(start + n, len(y))]
                if len(seg) < n:
                    seg = np.pad(seg, (0, n - len(seg)))
            else:
                seg = full[k]
            out.append(np.asarray(seg, dtype=np.float32))
        return out

    def _save_top_occlusion_patches_from_list(self, y, S, patch_importances, top_n, save_dir, file_name) -> Dict[str, dict]:
        base = Path(save_dir)
        groups = self.top_window_groups(patch_importances, top_n, file_name)
        for g, meta in groups.items():
            gdir = base / g
            gdir.mkdir(parents=True, exist_ok=True)
            if g != "all":
                for m, seg in zip(meta["windows"], self.window_audio(y, meta["windows"])):
                    name = (f"{file_name}__{g}{m['rank']}_patch_{m['type']}_{m['abs_importance']:.3f}_"
                            f"t{m['t_start']}-{m['t_end']}_f{m['f_start']}-{m['f_end']}.wav")
                    write_wav(gdir / name, seg, self.sr)
            with open(gdir / f"{file_name}__{g}_occlusion_patches_from_list.json", "w", encoding="utf-8") as f:
                json.dump(meta, f, indent=2, ensure_ascii=False)
        return groups

    # -- per-file / dataset drivers (host bookkeeping mirrors :808-1049; plots are out of scope) -----------
    def process_audio_file(self, audio_path: str, output_dir: Path, baseline_threshold: float = 0.3,
                           folder_name: str = "") -> Optional[Dict[str, Any]]:
        file_name = Path(audio_path).stem
        if self.checkpoint and str(audio_path) in self.checkpoint.load_processed_files():
            print("    Already processed, skipping...")
            return None
        result = self._compute_occlusion_map(audio_path=audio_path, baseline_threshold=baseline_threshold, verbose=True)
        if result.importance_map is None:
            if self.checkpoint:
                self.checkpoint.mark_as_processed(str(audio_path))
            return None
        rank, _ = dist.world()
        if rank == 0:
            track_dir = Path(output_dir) / folder_name / file_name if folder_name else Path(output_dir) / file_name
            track_dir.mkdir(parents=True, exist_ok=True)
            np.save(track_dir / f"saliency_{file_name}.npy", result.importance_map)   # raw map (reference keeps a PNG)
            self._save_top_occlusion_patches_from_list(result.y, result.S, result.patch_importances, self.top_n_windows,
                                                       track_dir / "top_windows", file_name)
            if self.checkpoint:
                self.checkpoint.mark_as_processed(str(audio_path))
        m = result.importance_map
        return {
            "file_path": str(audio_path), "file_name": file_name, "folder": folder_name, "method": self.method,
            "baseline_pred": float(result.baseline_pred), "mean_importance": float(m.mean()),
            "max_importance": float(m.max()), "min_importance": float(m.min()), "std_importance": float(m.std()),
            "p90_importance": float(np.percentile(m, 90)), "p10_importance": float(np.percentile(m, 10)),
        }

    def run_experiment(self, base_path, output_dir, models_to_process: Optional[list] = None,
                       max_samples_per_model: Optional[int] = None, baseline_threshold: float = 0.3,
                       resume: bool = True, results_path=None):
        import pandas as pd

        base_path, output_dir = Path(base_path), Path(output_dir)
        output_dir.mkdir(parents=True, exist_ok=True)
        results_path = Path(results_path) if results_path else output_dir / "spectrogram_explainability_results.json"
        saliency_dir = output_dir / "saliency_maps"
        saliency_dir.mkdir(parents=True, exist_ok=True)
        progress = output_dir / "spectrogram_results_progress.csv"
        results = pd.read_csv(progress).to_dict("records") if os.path.exists(progress) else []
        rank, _ = dist.world()
        for folder in sorted(base_path.iterdir()):
            if not folder.is_dir() or (models_to_process and folder.name not in models_to_process):
                continue
            files = sorted(list(folder.glob("*.mp3")) + list(folder.glob("*.wav")))
            if max_samples_per_model:
                files = files[:max_samples_per_model]
            for audio_file in files:
                try:
                    res = self.process_audio_file(str(audio_file), saliency_dir, baseline_threshold, folder.name)
                except AudioDecodeError as e:           # undecodable file: log it and go on (no checkpoint mark)
                    print(f"    Skipping {audio_file.name}: {e}")
                    continue
                if res:
                    results.append(res)
                    if rank == 0:
                        append_update_spectrogram_results({res["folder"]: {res["file_name"]: res}}, results_path)
                        pd.DataFrame(results).to_csv(progress, index=False)
        df = pd.DataFrame(results)
        if rank == 0 and len(df):
            df.to_csv(output_dir / f"spectrogram_results_{datetime.now().strftime('%Y%m%d_%H%M%S')}.csv", index=False)
        return df
