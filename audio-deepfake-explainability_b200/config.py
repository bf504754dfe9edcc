"""YAML -> constructor adapters: the reference's experiment runners, minus their plotting.

The reference maps its YAML files to constructor keywords inside two scripts:

* ``scripts/experiments/run_spectrogram_experiment.py:81-222``  (``spectrogram.* / explainability.* / dataset.* /
  output.* / checkpoint.*`` -> ``SpectrogramExplainability`` + ``run_experiment``)
* ``scripts/experiments/run_FBP_experiment.py:170-255``         (``bands.* / spectrogram.* / explainability.*`` ->
  ``FrequencyBandPerturbation`` + ``run_experiment``)

The functions below consume the SAME YAML structure with the SAME fallbacks - including the places where the runner's
fallback differs from the class default or from the shipped YAML (occlusion: ``n_mels 128`` / rise: ``n_mels 256``,
``patch_time_frames 2048``, ``patch_freq_percent 25``; FBP: ``transition.hz 200``, ``min_hz 20``, ``max_hz 2000``,
``n_iter 32``, ``presets {}`` which makes a missing ``presets`` block select the built-in *default* bank whatever
``preset`` says) - so a reference config file drives the B200 engine unchanged.  The predictor is built from
``model.*``: ``model.local_model`` must be a local checkpoint path (no network here) or the literal ``random-init[:seed]``.
"""
from __future__ import annotations

from pathlib import Path
from typing import Any, Dict, Optional, Union

ConfigLike = Union[str, Path, Dict[str, Any]]


def load_yaml(path) -> Dict[str, Any]:
    """``load_yaml`` of both runners (run_spectrogram_experiment.py:34-37): ``yaml.safe_load`` of a UTF-8 file."""
    import yaml

    with open(path, "r", encoding="utf-8") as f:
        return yaml.safe_load(f) or {}


def _as_config(config: ConfigLike) -> Dict[str, Any]:
    return config if isinstance(config, dict) else load_yaml(config)


def predictor_from_config(model_cfg: Optional[Dict[str, Any]], **engine_kw):
    """``model.*`` -> ``B200Predictor`` (runner: run_spectrogram_experiment.py:120-155, run_FBP_experiment.py:93-129).
    The remote Gradio predictor (``local: false``) is outside the hot path: the B200 engine is always local."""
    from .sonics_api import B200Predictor

    model_cfg = model_cfg or {}
    name = str(model_cfg.get("local_model", "awsaf49/sonics-spectttra-alpha-120s"))
    device = model_cfg.get("device", "cuda")
    if device != "cuda":
        raise RuntimeError(f"model.device={device!r}: the B200 engine has no CPU path")
    if name.startswith("random-init"):
        seed = int(name.split(":", 1)[1]) if ":" in name else 0
        return B200Predictor.random_init(seed, **engine_kw)
    return B200Predictor.from_pretrained(name, device=device, **engine_kw)


def spectrogram_explainer_kwargs(config: ConfigLike, checkpoint_dir=None) -> Dict[str, Any]:
    """Constructor keywords exactly as run_spectrogram_experiment.py:157-205 derives them."""
    config = _as_config(config)
    s = config.get("spectrogram", {}) or {}
    e = config.get("explainability", {}) or {}
    viz = e.get("visualization", {}) or {}
    method = e.get("method", "rise")
    common = dict(
        sr=s.get("sr", 44100), duration=s.get("duration", 120), n_fft=s.get("n_fft", 2048),
        hop_length=s.get("hop_length", 512), win_length=s.get("win_length", 2048), n_iter=s.get("n_iter", 256),
        spec_type=s.get("spec_type", "mel"), fmax=s.get("fmax", None), checkpoint_dir=checkpoint_dir,
        highlight_percent=viz.get("highlight_percent", 20.0), abs_threshold=viz.get("abs_threshold", None),
    )
    if method == "rise":
        r = e.get("rise", {}) or {}
        return dict(common, n_mels=s.get("n_mels", 256), method="rise", use_original_audio=False,
                    n_masks=r.get("n_masks", 500), mask_probability=r.get("mask_probability", 0.5))
    o = e.get("occlusion", {}) or {}
    return dict(common, n_mels=s.get("n_mels", 128), top_n_windows=o.get("top_n_windows", 5), method="occlusion",
                use_original_audio=o.get("use_original_audio", True),
                patch_time_frames=o.get("patch_time_frames", 2048), stride_time_frames=o.get("stride_time_frames", 2048),
                patch_freq_percent=o.get("patch_freq_percent", 25.0), stride_freq_percent=o.get("stride_freq_percent", 25.0))


def fbp_kwargs(config: ConfigLike, checkpoint_dir=None, save_fbp_audio: str = "none") -> Dict[str, Any]:
    """Constructor keywords exactly as run_FBP_experiment.py:222-253 derives them (``save_fbp_audio`` is the runner's
    ``--save-fbp-audio`` flag: none | separated | reversed)."""
    config = _as_config(config)
    b = config.get("bands", {}) or {}
    t = b.get("transition", {}) or {}
    s = config.get("spectrogram", {}) or {}
    e = config.get("explainability", {}) or {}
    return dict(
        preset=b.get("preset", "default"), presets=b.get("presets", {}), attenuation=float(b.get("attenuation", 0.0)),
        transition_mode=str(t.get("mode", "rel")), transition_hz=float(t.get("hz", 200.0)),
        transition_rel=float(t.get("rel", 0.2)), transition_min_hz=float(t.get("min_hz", 20.0)),
        transition_max_hz=float(t.get("max_hz", 2000.0)),
        sr=int(s.get("sr", 44100)), duration=int(s.get("duration", 120)), n_mels=int(s.get("n_mels", 128)),
        n_fft=int(s.get("n_fft", 2048)), hop_length=int(s.get("hop_length", 512)), win_length=int(s.get("win_length", 2048)),
        n_iter=int(s.get("n_iter", 32)), spec_type=str(s.get("spec_type", "stft")), fmax=s.get("fmax", None),
        use_original_audio=bool(e.get("use_original_audio", False)), use_separation=bool(e.get("use_separation", False)),
        separation_model=str(e.get("separation_model", "spleeter:2stems")),
        separation_targets=tuple(e.get("separation_targets", ("vocals0", "accompaniment0"))),
        normalize_loudness=bool(e.get("normalize_loudness", True)), lufs=float(e.get("lufs", -14.0)),
        checkpoint_dir=checkpoint_dir, save_perturbed_audio_only=save_fbp_audio == "separated",
        save_reversed_perturbed_audio_only=save_fbp_audio == "reversed",
    )


def _output_dir(config: Dict[str, Any], default_name: str) -> Path:
    out = config.get("output", {}) or {}
    return Path(out.get("result_path")) / str(out.get("experiment_name", default_name))


def _checkpoint_dir(config: Dict[str, Any], output_dir: Path, no_checkpoint: bool) -> Optional[Path]:
    if (config.get("checkpoint", {}) or {}).get("enabled", True) and not no_checkpoint:
        d = output_dir / "checkpoints"
        d.mkdir(parents=True, exist_ok=True)
        return d
    return None


def run_spectrogram_experiment(config: ConfigLike, predictor=None, no_checkpoint: bool = False, resume: bool = False):
    """``main()`` of run_spectrogram_experiment.py without the plotting: build predictor + explainer from the YAML and call
    ``run_experiment`` with the runner's arguments (:207-217).  Returns the results DataFrame."""
    from .spectrogram_explainability import SpectrogramExplainability

    config = _as_config(config)
    d = config.get("dataset", {}) or {}
    e = config.get("explainability", {}) or {}
    output_dir = _output_dir(config, "spectrogram_exp")
    output_dir.mkdir(parents=True, exist_ok=True)
    ckpt = _checkpoint_dir(config, output_dir, no_checkpoint)
    predictor = predictor if predictor is not None else predictor_from_config(config.get("model"))
    explainer = SpectrogramExplainability(predictor=predictor, **spectrogram_explainer_kwargs(config, ckpt))
    return explainer.run_experiment(
        base_path=Path(d.get("base_path")), output_dir=output_dir, models_to_process=d.get("models_to_process"),
        max_samples_per_model=d.get("max_samples_per_model"), baseline_threshold=e.get("baseline_threshold", 0.3),
        resume=resume or (not no_checkpoint), results_path=output_dir / f"spectrogram_results_{e.get('method', 'rise')}.json")


def run_fbp_experiment(config: ConfigLike, predictor=None, no_checkpoint: bool = False, save_fbp_audio: str = "none"):
    """``main()`` of run_FBP_experiment.py without the plotting (:214-262)."""
    from .dsp_band_ops import FrequencyBandPerturbation

    config = _as_config(config)
    d = config.get("dataset", {}) or {}
    output_dir = _output_dir(config, "exp")
    output_dir.mkdir(parents=True, exist_ok=True)
    ckpt = _checkpoint_dir(config, output_dir, no_checkpoint)
    predictor = predictor if predictor is not None else predictor_from_config(config.get("model"))
    fbp = FrequencyBandPerturbation(predictor=predictor, **fbp_kwargs(config, ckpt, save_fbp_audio))
    return fbp.run_experiment(base_path=Path(d.get("base_path")), output_dir=output_dir,
                              models_to_process=d.get("models_to_process"),
                              max_samples_per_model=d.get("max_samples_per_model"), results_path=output_dir / "fbp_results.json")
