"""Frequency-band perturbation - drop-in for ``FrequencyBandPerturbation`` (src/dsp_band_ops.py:303-1008).

Same constructor keywords, band presets, transition rules, ``FBDResult`` and ``<file>_bands_metadata.json`` layout.
The reference's serial band loop (:573-653: attenuate a band on |S|, iSTFT, optional RMS match, predict) becomes one
batched device sweep over all bands; the per-bin gain vectors ``keep + attenuation * (1 - keep)`` are tiny float64
host arithmetic (grid.band_gain_table) and are applied inside the iSTFT load stage on the GPU.
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import Sequence, Any, Dict, List, NamedTuple, Optional, Tuple

import numpy as np

from . import dist, grid, mel_host
from .audio_io import AudioDecodeError, load_audio, write_wav
from .grid import FREQUENCY_BAND_PRESETS, smooth_band_keep_mask  # noqa: F401  (re-exported like the reference module)
from .sonics_api import B200Predictor
from .spectrogram_explainability import amplitude_to_db_refmax


def match_rms(ref: np.ndarray, x: np.ndarray, eps: float = 1e-8) -> np.ndarray:
    """Host mirror of src/dsp_band_ops.py:228-233 (the sweep applies the same gain on the GPU)."""
    r_ref = float(np.sqrt(np.mean(ref ** 2) + eps))
    r_x = float(np.sqrt(np.mean(x ** 2) + eps))
    return x if r_x < eps else x * (r_ref / r_x)


class FBDResult(NamedTuple):
    importance_map: Optional[np.ndarray]
    spectrogram_db: np.ndarray
    baseline_pred: float
    y: np.ndarray
    S: np.ndarray
    batch_importances: Optional[list]


class FrequencyBandPerturbation:
    def __init__(self, predictor, preset: str = "default", presets: Optional[Dict[str, List[Tuple[int, int]]]] = None,
                 attenuation: float = 0.0, transition_mode: str = "rel", transition_hz: float = 0.0,
                 transition_rel: float = 0.0, transition_min_hz: float = 0.0, transition_max_hz: float = 0.0,
                 sr: int = 44100, duration: int = 120, n_mels: int = 128, n_fft: int = 2048, hop_length: int = 512,
                 win_length: int = 2048, n_iter: int = 256, spec_type: str = "stft", fmax: Optional[float] = None,
                 use_original_audio: bool = False, use_separation: bool = False, separation_model: str = "spleeter:2stems",
                 separation_targets: Tuple[str, ...] = ("vocals0", "accompaniment0"), normalize_loudness: bool = True,
                 lufs: Optional[float] = None, checkpoint_dir=None, save_perturbed_audio_only: bool = False,
                 save_reversed_perturbed_audio_only: bool = False, mel_seed: int = 0, nnls_iter: int = 16):
        if not isinstance(predictor, B200Predictor):
            raise TypeError(f"the B200 band sweep needs a B200Predictor, got {type(predictor).__name__}")
        self.predictor = predictor
        self.preset, self.presets = preset, presets
        table = presets if presets is not None else FREQUENCY_BAND_PRESETS
        self.bands = [tuple(b) for b in table.get(preset, FREQUENCY_BAND_PRESETS["default"])]
        self.attenuation = attenuation
        self.transition_mode, self.transition_hz, self.transition_rel = transition_mode, transition_hz, transition_rel
        self.transition_min_hz, self.transition_max_hz = transition_min_hz, transition_max_hz
        self.sr, self.duration, self.n_mels = sr, duration, n_mels
        self.n_fft, self.hop_length, self.win_length, self.n_iter = n_fft, hop_length, win_length, n_iter
        self.spec_type = spec_type.lower()
        # The reference supports only 'stft' here (:357-359).  'mel' is a BUILDER-DEFINED extension: the band
        # gain is evaluated at the centre frequency of every mel bin, applied to the power mel spectrogram and inverted with
        # the seeded NNLS + Griffin-Lim of the mel occlusion variant (oracle/mel.py); no reference parity exists for it.
        if self.spec_type not in ("stft", "mel"):
            raise ValueError("FrequencyBandPerturbation supports spec_type='stft' (reference) or 'mel' (builder-defined)")
        self.mel_seed, self.nnls_iter = int(mel_seed), int(nnls_iter)
        if (n_fft, hop_length, win_length) != (2048, 512, 2048):
            raise NotImplementedError("the CUDA STFT/iSTFT kernels are built for n_fft=2048, hop=512, win=2048")
        self.fmax = fmax if fmax is not None else sr // 2
        self.use_original_audio = use_original_audio
        if use_separation:
            raise NotImplementedError("use_separation needs Spleeter (TensorFlow), which is outside the hot path; "
                                      "pass separated stems to compute_component_importance() instead")
        self.use_separation, self.separation_model, self.separation_targets = use_separation, separation_model, separation_targets
        self.normalize_loudness, self.lufs = normalize_loudness, lufs
        self.save_perturbed_audio_only = save_perturbed_audio_only
        self.save_reversed_perturbed_audio_only = save_reversed_perturbed_audio_only
        self.checkpoint_dir = Path(checkpoint_dir) if checkpoint_dir else None

    @classmethod
    def from_config(cls, config, predictor, checkpoint_dir=None, save_fbp_audio: str = "none") -> "FrequencyBandPerturbation":
        """Explainer from a reference YAML file / dict, keys and fallbacks as run_FBP_experiment.py:222-253."""
        from .config import fbp_kwargs
        return cls(predictor=predictor, **fbp_kwargs(config, checkpoint_dir, save_fbp_audio))

    def _band_transition_width(self, low: float, high: float) -> float:
        return grid.band_transition_width(low, high, self.transition_mode, self.transition_rel, self.transition_min_hz,
                                          self.transition_max_hz, self.transition_hz)

    def band_gains(self) -> np.ndarray:
        """float64 ``[n_bands, n_freq]`` = keep + attenuation * (1 - keep)  (:574-576)."""
        return grid.band_gain_table(self.bands, self.sr, self.n_fft, self.attenuation, self.transition_mode,
                                    self.transition_rel, self.transition_min_hz, self.transition_max_hz, self.transition_hz)

    def _predict(self, wave: np.ndarray) -> float:
        return float(self.predictor.predict(wave, self.sr))

    # -- the hot path ---------------------------------------------------------------------------------
    def _compute_component_importance(self, sig: np.ndarray, component_name: str, audio_path: str = "",
                                      audio_root: Optional[Path] = None, file_name: Optional[str] = None,
                                      **_ignored) -> Optional[FBDResult]:
        eng = self.predictor.engine
        sig = np.ascontiguousarray(np.asarray(sig, dtype=np.float32))
        if self.spec_type == "mel":
            return self._mel_component_importance(sig, component_name)
        gains = self.band_gains()
        _, world = dist.world()
        if world == 1 and len(self.bands) <= eng.copies_per_chunk:
            # baseline and band copies in one device pass (the baseline rides in the band copies' forward when it can)
            base, probs2 = eng.fbp_sweep_tracks(sig[None, :], gains.astype(np.float32), self.normalize_loudness)
            orig_prob, probs = float(base[0]), probs2[0]
            S = eng.spectrogram()                      # the swept track is the engine's current track
        else:
            eng.set_track(sig)
            orig_prob = float(eng.predict_track())
            S = eng.spectrogram()
            probs = dist.sharded_sweep(lambda g: eng.fbp_sweep(g, self.normalize_loudness), gains.astype(np.float32))
        deltas = [float(orig_prob - float(p)) for p in probs]
        if (self.save_perturbed_audio_only or self.save_reversed_perturbed_audio_only) and audio_root is not None:
            self._save_band_audio(sig, gains, deltas, Path(audio_root), component_name, file_name or "track")
            return None
        batch = [{"component": component_name, "low": float(lo), "high": float(hi), "importance": d}
                 for (lo, hi), d in zip(self.bands, deltas)]
        rows = grid.band_bin_ranges(self.bands, self.sr, self.n_fft)
        importance_map = eng.band_map(rows, np.asarray(deltas, dtype=np.float64))
        return FBDResult(importance_map, amplitude_to_db_refmax(S), orig_prob, sig, S, batch)

    def _mel_component_importance(self, sig: np.ndarray, component_name: str) -> FBDResult:
        """FBP over the mel spectrogram (builder-defined): per band, mel bins are scaled by the band gain at their centre
        frequency, the result is inverted (NNLS -> Griffin-Lim) and classified; ``importance_map`` is ``[n_mels, n_time]``."""
        eng = self.predictor.engine
        eng.set_track(sig)
        basis = mel_host.mel_filterbank(self.sr, self.n_fft, self.n_mels, 0.0, None)
        pinv, step = mel_host.nnls_operators(basis)
        eng.set_mel_basis(basis, pinv, step)
        S = eng.mel_spectrogram()
        gains = mel_host.mel_band_gain_table(self.bands, self.sr, self.n_mels, self.attenuation, self.transition_mode,
                                             self.transition_rel, self.transition_min_hz, self.transition_max_hz, self.transition_hz)
        orig_prob = float(eng.predict_track())
        rank, world = dist.world()
        lo, hi = grid.shard_range(len(gains), rank, world)
        local = (eng.mel_sweep(eng.MASK_BAND_GAIN, gains[lo:hi].astype(np.float32), self.n_iter, self.nnls_iter, self.mel_seed,
                               first_index=lo) if hi > lo else np.zeros(0, np.float32))
        probs = dist.gather_shards(local, len(gains))
        deltas = [float(orig_prob - float(p)) for p in probs]
        batch = [{"component": component_name, "low": float(lo_), "high": float(hi_), "importance": d}
                 for (lo_, hi_), d in zip(self.bands, deltas)]
        importance_map = np.zeros(S.shape, dtype=np.float64)
        for rows, d in zip(mel_host.mel_band_rows(self.bands, self.sr, self.n_mels), deltas):
            importance_map[rows, :] += d
        return FBDResult(importance_map, mel_host.power_to_db_refmax(S), orig_prob, sig, S, batch)

    def compute_importance_batch(self, signals: Sequence[np.ndarray], component_name: str = "mixture",
                                 materialize_maps: bool = False) -> List[FBDResult]:
        """``_compute_component_importance`` for a batch of equal-length signals in shared launches (BASELINE configs[2]:
        64 tracks x the high_resolution bank).  The reference handles one file at a time (:529-666); here the band copies
        of as many tracks as fit a chunk go through one iSTFT launch and one classifier forward, with results identical
        to the per-track call.  ``S`` / ``spectrogram_db`` are not materialised (30 MB per track on the host) - ask the
        per-track method for them.  ``importance_map`` is a read-only broadcast view of the map's single distinct column
        (``grid.band_map_view``; same values, same shape and dtype as the per-track method's array - a third of the batch's
        wall time was the device -> host copy of 64 x 30.8 MB of repeated columns); ``materialize_maps=True`` returns the
        device-built writable arrays instead."""
        if not signals:
            return []
        waves = np.stack([np.ascontiguousarray(np.asarray(s, dtype=np.float32)) for s in signals])
        eng = self.predictor.engine
        gains = self.band_gains()
        base, probs = eng.fbp_sweep_tracks(waves, gains.astype(np.float32), self.normalize_loudness)
        rows = grid.band_bin_ranges(self.bands, self.sr, self.n_fft)
        out = []
        for i in range(len(signals)):
            orig_prob = float(base[i])
            deltas = [float(orig_prob - float(p)) for p in probs[i]]
            batch = [{"component": component_name, "low": float(lo), "high": float(hi), "importance": d}
                     for (lo, hi), d in zip(self.bands, deltas)]
            if materialize_maps:
                importance_map = eng.band_map(rows, np.asarray(deltas, dtype=np.float64))
            else:
                importance_map = grid.band_map_view(rows, deltas, self.n_fft // 2 + 1, 1 + waves.shape[1] // self.hop_length)
            out.append(FBDResult(importance_map, None, orig_prob, waves[i], None, batch))
        return out

    def _save_band_audio(self, sig, gains, deltas, audio_root: Path, component: str, file_name: str) -> None:
        """separated_bands / reversed_separated_bands WAV layout (:608-639)."""
        eng = self.predictor.engine
        sub = "separated_bands" if self.save_perturbed_audio_only else "reversed_separated_bands"
        out_dir = audio_root / component / sub / "freq_batches"
        out_dir.mkdir(parents=True, exist_ok=True)
        g = (1.0 - gains) if self.save_perturbed_audio_only else gains
        audio = eng.band_audio(g.astype(np.float32))
        for (lo, hi), d, y_b in zip(self.bands, deltas, audio):
            y_b = y_b.astype(np.float64)
            if self.normalize_loudness:
                y_b = match_rms(sig.astype(np.float64), y_b)
            peak = np.max(np.abs(y_b))
            y_out = y_b / peak * 0.99 if peak > 0 else y_b
            name = f"{file_name}__{component}__{int(lo)}-{int(hi)}Hz_{grid.importance_type(d)}_{d:+.3f}.wav"
            write_wav(out_dir / name, y_out.astype(np.float32), self.sr)

    def _compute_importance(self, audio_path: str, track_output_dir: Optional[Path] = None, file_name: Optional[str] = None,
                            **_ignored) -> list:
        y, _ = load_audio(audio_path, sr=self.sr, duration=self.duration, mono=True, resample=self.predictor.engine.resample)
        audio_only = self.save_perturbed_audio_only or self.save_reversed_perturbed_audio_only
        res = self._compute_component_importance(y, "mixture", audio_path, audio_root=track_output_dir if audio_only else None,
                                                 file_name=file_name)
        return [res] if res is not None else []

    def _save_frequency_band_importances(self, batch_importances: list, file_name: str, save_dir: Path) -> dict:
        save_dir.mkdir(parents=True, exist_ok=True)
        meta = {"file_name": file_name, "bands": []}
        for p in batch_importances:
            imp = p["importance"]
            meta["bands"].append({"component": p.get("component", "mixture"), "low": p["low"], "high": p["high"],
                                  "importance": imp, "abs_importance": abs(imp), "type": grid.importance_type(imp)})
        with open(save_dir / f"{file_name}_bands_metadata.json", "w", encoding="utf-8") as f:
            json.dump(meta, f, indent=2, ensure_ascii=False)
        return meta

    def process_audio_file(self, audio_path: str, output_dir: Path, folder_name: str = "", **_ignored) -> Optional[Dict[str, Any]]:
        file_name = Path(audio_path).stem
        track_dir = Path(output_dir) / folder_name / file_name if folder_name else Path(output_dir) / file_name
        track_dir.mkdir(parents=True, exist_ok=True)
        results = self._compute_importance(audio_path, track_output_dir=track_dir, file_name=file_name)
        if not results:
            return None
        rank, _ = dist.world()
        summary = {}
        total = None
        for r in results:
            comp = r.batch_importances[0]["component"] if r.batch_importances else "mixture"
            if rank == 0:
                self._save_frequency_band_importances(r.batch_importances, file_name, track_dir / comp)
                np.save(track_dir / comp / f"fbp_saliency_{file_name}.npy", r.importance_map)
            m = r.importance_map
            summary[comp] = {"baseline_pred_mean": float(r.baseline_pred), "mean_importance": float(m.mean()),
                             "max_importance": float(m.max()), "min_importance": float(m.min()), "std_importance": float(m.std())}
            total = m if total is None else total + m
        return {"file_path": str(audio_path), "file_name": file_name, "folder": folder_name, "components": summary,
                "global_mean_importance": float(total.mean()), "global_max_importance": float(total.max()),
                "global_min_importance": float(total.min()), "global_std_importance": float(total.std())}

    def run_experiment(self, base_path, output_dir, models_to_process: Optional[list] = None,
                       max_samples_per_model: Optional[int] = None, **_ignored):
        import pandas as pd

        base_path, output_dir = Path(base_path), Path(output_dir)
        bands_dir = output_dir / "bands"
        bands_dir.mkdir(parents=True, exist_ok=True)
        rows = []
        for folder in sorted(base_path.iterdir()):
            if not folder.is_dir() or (models_to_process and folder.name not in models_to_process):
                continue
            files = sorted(list(folder.glob("*.mp3")) + list(folder.glob("*.wav")))
            if max_samples_per_model:
                files = files[:max_samples_per_model]
            for audio_file in files:
                try:
                    res = self.process_audio_file(str(audio_file), bands_dir, folder.name)
                except AudioDecodeError as e:           # undecodable file: log it and go on (no checkpoint mark)
                    print(f"    Skipping {audio_file.name}: {e}")
                    continue
                if res:
                    flat = {k: v for k, v in res.items() if k != "components"}
                    rows.append(flat)
        return pd.DataFrame(rows)
