"""Predictor boundary - B200-native mirror of ``LocalSonnics`` (src/sonics_api.py:230-317).

Same duck type the reference's explainers consume (``predict(wave, sr) -> float`` in [0, 1] = P(fake),
``predict_from_file``, ``predict_batch_from_files``, ``from_pretrained``) plus the batched sweeps that replace the
reference's one-evaluation-per-call loops.  Everything computes on the GPU through ``libb200xai.so``;
there is no CPU or eager-PyTorch fallback, and an error is raised rather than turned into a 0.0 prediction
(contrast src/spectrogram_explainability.py:357-362).
"""
from __future__ import annotations

import os
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Union

import numpy as np

from .audio_io import load_audio
from .engine import Engine
from .weights import ALPHA_120S, SpecTTTraConfig, random_state_dict


def _load_state_dict_file(path: Path) -> Dict[str, np.ndarray]:
    if path.suffix == ".npz":
        with np.load(path) as z:
            return {k: z[k] for k in z.files}
    import torch

    sd = torch.load(str(path), map_location="cpu")
    if isinstance(sd, dict) and "state_dict" in sd:
        sd = sd["state_dict"]
    return {k: v.detach().float().numpy() for k, v in sd.items() if hasattr(v, "detach")}


class B200Predictor:
    """Local SpecTTTra predictor running on one B200."""

    def __init__(self, model_name: str = "awsaf49/sonics-spectttra-alpha-120s", device: str = "cuda",
                 state_dict: Optional[Dict[str, np.ndarray]] = None, cfg: SpecTTTraConfig = ALPHA_120S,
                 copies_per_chunk: int = 128, max_samples: int = 120 * 44100, device_index: Optional[int] = None):
        if device != "cuda":
            raise RuntimeError(f"B200Predictor runs on CUDA only (device={device!r}); it has no CPU path")
        self.model_name = model_name
        self.device = device
        self.cfg = cfg
        if device_index is None:
            device_index = int(os.environ.get("LOCAL_RANK", "0"))
        if state_dict is None:
            raise RuntimeError(
                "no weights given: there is no network in this environment, so pass state_dict=... , use "
                "B200Predictor.from_pretrained(<local .npz/.pt path>) or B200Predictor.random_init(seed)")
        self.engine = Engine(cfg, state_dict, copies_per_chunk=copies_per_chunk, max_samples=max_samples, device=device_index)

    # -- constructors ----------------------------------------------------------------------------
    @classmethod
    def from_pretrained(cls, name: str, device: str = "cuda", **kw) -> "B200Predictor":
        """``name``: local checkpoint path (sonics state-dict as .npz / torch .pt). HF hub ids need a network."""
        p = Path(name)
        if not p.exists():
            raise RuntimeError(f"checkpoint '{name}' not found locally; HF hub download is unavailable (no network). "
                               "Use B200Predictor.random_init(seed) for synthetic-weight runs.")
        return cls(model_name=name, device=device, state_dict=_load_state_dict_file(p), **kw)

    @classmethod
    def random_init(cls, seed: int = 0, cfg: SpecTTTraConfig = ALPHA_120S, head_gain: float = 1.0, **kw) -> "B200Predictor":
        return cls(model_name=f"random-init-seed{seed}", state_dict=random_state_dict(cfg, seed, head_gain), cfg=cfg, **kw)

    # -- LocalSonnics surface (src/sonics_api.py:259-317) -----------------------------------------
    def predict(self, audio_wave: np.ndarray, sr: int) -> float:
        """Fake probability of one mono waveform.  ``sr`` is ignored, exactly as in the reference (:268-271)."""
        return float(self.engine.predict(np.asarray(audio_wave, dtype=np.float32)))

    def predict_from_file(self, audio_path: Union[str, Path], sr: int = 44100, duration: Optional[float] = None) -> float:
        y, _ = load_audio(str(audio_path), sr=sr, duration=duration, mono=True, resample=self.engine.resample)
        return self.predict(y, sr)

    def predict_batch_from_files(self, audio_paths: Sequence[Union[str, Path]], sr: int = 44100,
                                 duration: Optional[float] = None, verbose: bool = True, **kwargs) -> List[float]:
        probs = []
        for idx, path in enumerate(audio_paths):
            if verbose:
                print(f"   Predicting {idx + 1}/{len(audio_paths)}: {Path(path).name}")
            probs.append(self.predict_from_file(path, sr=sr, duration=duration))
            if verbose:
                print(f"      -> Fake prob: {probs[-1]:.4f}")
        return probs

    # -- batched natives ---------------------------------------------------------------------------
    def predict_batch(self, waves: np.ndarray) -> np.ndarray:
        """Probabilities of ``[count, L]`` equal-length waves in one device pass."""
        return self.engine.predict(np.asarray(waves, dtype=np.float32))

    def stem_mask_sweep(self, stems: np.ndarray, masks: np.ndarray) -> np.ndarray:
        """AudioLIME recombinations: ``[N, 2] = (1 - p, p)`` like ``predict_fn_unified`` (src/lime_explainer.py:283-301)."""
        p = self.engine.stem_sweep(stems, masks).astype(np.float64)
        return np.stack([1.0 - p, p], axis=1)

    def close(self) -> None:
        self.engine.close()


def predict_from_file(predictor, audio_path, **kwargs) -> float:
    """Module-level dispatcher kept for drop-in compatibility (src/sonics_api.py:319-331)."""
    return predictor.predict_from_file(audio_path, **kwargs)


def predict_batch_from_files(predictor, audio_paths, verbose: bool = True, **kwargs) -> List[float]:
    """src/sonics_api.py:333-345."""
    return predictor.predict_batch_from_files(audio_paths, verbose=verbose, **kwargs)
