"""YAML -> constructor adapters against VERBATIM copies of the reference's two shipped experiment configs
(tests/golden/ref_spectrogram_explainability.yaml = configs/Spec_occlusion_configs/spectrogram_explainability.yaml,
tests/golden/ref_fbp_experiment.yaml = configs/FBP_configs/fbp_experiment.yaml) and against the runner fallbacks
(scripts/experiments/run_spectrogram_experiment.py:157-205, run_FBP_experiment.py:222-253)."""
from pathlib import Path

import numpy as np

from audio_deepfake_explainability_b200 import config as cfgmod
from audio_deepfake_explainability_b200 import grid
from audio_deepfake_explainability_b200.dsp_band_ops import FREQUENCY_BAND_PRESETS, FrequencyBandPerturbation
from audio_deepfake_explainability_b200.sonics_api import B200Predictor
from audio_deepfake_explainability_b200.spectrogram_explainability import SpectrogramExplainability

GOLDEN = Path(__file__).parent / "golden"


def _predictor():
    return object.__new__(B200Predictor)            # the constructors only type-check the predictor; no GPU needed


def test_shipped_occlusion_yaml_maps_like_the_runner():
    path = GOLDEN / "ref_spectrogram_explainability.yaml"
    kw = cfgmod.spectrogram_explainer_kwargs(path, checkpoint_dir=None)
    assert kw == dict(sr=44100, duration=120, n_fft=2048, hop_length=512, win_length=2048, n_mels=512, n_iter=256,
                      spec_type="stft", fmax=None, top_n_windows=5, method="occlusion", use_original_audio=False,
                      patch_time_frames=1024, stride_time_frames=1024, patch_freq_percent=20.0, stride_freq_percent=10.0,
                      checkpoint_dir=None, highlight_percent=25.0, abs_threshold=None)
    ex = SpectrogramExplainability.from_config(path, _predictor())
    assert (ex.method, ex.spec_type, ex.sr, ex.top_n_windows, ex.use_original_audio) == ("occlusion", "stft", 44100, 5, False)
    # the shipped configuration: 44.1 kHz x 120 s -> n_time 10 336; 1024 x 20 % at 10 % stride -> 10 x 9 = 90 windows
    n_freq, n_time = grid.stft_shape(44100 * 120, ex.n_fft, ex.hop_length)
    wins = grid.occlusion_windows(n_freq, n_time, ex.patch_time_frames, ex.stride_time_frames, ex.patch_freq_percent,
                                  ex.stride_freq_percent)
    assert (n_freq, n_time, len(wins)) == (1025, 10336, 90)
    assert cfgmod.load_yaml(path)["explainability"]["baseline_threshold"] == 0.00001


def test_occlusion_runner_fallbacks_differ_from_class_defaults():
    kw = cfgmod.spectrogram_explainer_kwargs({"explainability": {"method": "occlusion"}})
    assert (kw["n_mels"], kw["patch_time_frames"], kw["stride_time_frames"], kw["patch_freq_percent"], kw["stride_freq_percent"]) == \
        (128, 2048, 2048, 25.0, 25.0)
    assert kw["spec_type"] == "mel" and kw["use_original_audio"] is True and kw["highlight_percent"] == 20.0
    rise = cfgmod.spectrogram_explainer_kwargs({})           # method falls back to 'rise' (:158)
    assert (rise["method"], rise["n_mels"], rise["n_masks"], rise["mask_probability"], rise["use_original_audio"]) == \
        ("rise", 256, 500, 0.5, False)
    assert "top_n_windows" not in rise


def test_shipped_fbp_yaml_maps_like_the_runner():
    path = GOLDEN / "ref_fbp_experiment.yaml"
    kw = cfgmod.fbp_kwargs(path)
    assert kw["preset"] == "default" and kw["attenuation"] == 0.25
    assert (kw["transition_mode"], kw["transition_rel"], kw["transition_min_hz"], kw["transition_max_hz"], kw["transition_hz"]) == \
        ("rel", 0.2, 5.0, 500.0, 200.0)
    assert (kw["sr"], kw["n_mels"], kw["n_iter"], kw["spec_type"]) == (44100, 512, 256, "stft")
    assert kw["normalize_loudness"] is False and kw["lufs"] == -14.0 and kw["use_separation"] is False
    assert kw["separation_targets"] == ("vocals0", "drums0", "bass0", "other0")
    fbp = FrequencyBandPerturbation.from_config(path, _predictor())
    assert fbp.bands == [(20, 100), (100, 250), (250, 2000), (2000, 4000), (4000, 8000), (8000, 16000)]
    # the YAML's preset table equals the module constant (dsp_band_ops.py:212-226)
    table = cfgmod.load_yaml(path)["bands"]["presets"]
    assert {k: [tuple(b) for b in v] for k, v in table.items()} == {k: [tuple(b) for b in v] for k, v in FREQUENCY_BAND_PRESETS.items()}
    cfg = cfgmod.load_yaml(path)
    cfg["bands"]["preset"] = "high_resolution"
    hi = FrequencyBandPerturbation.from_config(cfg, _predictor())
    assert len(hi.bands) == 13 and hi.bands[-1] == (16000, 21000)
    widths = [hi._band_transition_width(lo, hi_) for lo, hi_ in hi.bands]
    assert widths == [8.0, 8.0, 30.0, 50.0, 100.0, 200.0, 400.0, 400.0, 400.0, 400.0, 400.0, 500.0, 500.0]
    assert hi.band_gains().shape == (13, 1025) and np.isclose(hi.band_gains().min(), 0.25)


def test_fbp_runner_fallbacks():
    kw = cfgmod.fbp_kwargs({})
    assert (kw["transition_min_hz"], kw["transition_max_hz"], kw["transition_hz"], kw["n_iter"], kw["attenuation"]) == \
        (20.0, 2000.0, 200.0, 32, 0.0)
    assert kw["presets"] == {} and kw["normalize_loudness"] is True
    # presets={} (a YAML without a presets block) selects the built-in default bank whatever `preset` says (:337-340)
    fbp = FrequencyBandPerturbation.from_config({"bands": {"preset": "high_resolution"}}, _predictor())
    assert fbp.bands == [tuple(b) for b in FREQUENCY_BAND_PRESETS["default"]]
    kw = cfgmod.fbp_kwargs({}, save_fbp_audio="reversed")
    assert kw["save_reversed_perturbed_audio_only"] and not kw["save_perturbed_audio_only"]
