"""ORACLE tooling (test infrastructure only): generate golden vectors from the REFERENCE ITSELF.

Run in the build container (``python -m oracle.make_golden_from_reference``); /root/reference does not
exist on the GPU box, so the outputs are committed under tests/golden/.

The reference modules ``src/spectrogram_explainability.py`` and ``src/dsp_band_ops.py`` cannot be
imported as-is here (top-level ``import librosa / soundfile / matplotlib / seaborn``, none installed).
This script installs stub modules for exactly those missing third-party imports - ``librosa.{load, stft,
istft, magphase, fft_frequencies, amplitude_to_db}`` are served by the restatements in ``oracle/dsp.py``,
``soundfile.write`` captures the arrays, matplotlib/seaborn are inert - and then imports and runs the
reference's *own, unmodified* functions:

  * ``SpectrogramExplainability._compute_occlusion_map`` and ``_save_top_occlusion_patches_from_list``
  * ``FrequencyBandPerturbation._compute_component_importance``; ``smooth_band_keep_mask``; ``match_rms``;
    ``FREQUENCY_BAND_PRESETS``; ``_band_transition_width``

with a cheap deterministic stub predictor.  What this pins: window/band indexing, iteration order,
map accumulation and normalisation, top-k grouping/ordering/JSON layout, keep-mask values, RMS matching.
What it cannot pin: librosa's and sonics' own numerics (absent third-party packages).
"""
from __future__ import annotations

import json
import sys
import tempfile
import types
from pathlib import Path
from unittest import mock

import numpy as np

REPO = Path(__file__).resolve().parents[1]
REF = Path("/root/reference")
GOLDEN = REPO / "tests" / "golden"


class EnergyPredictor:
    """Deterministic stand-in classifier: sigmoid of a weighted log band-energy ratio."""

    def __init__(self, sr: int):
        self.sr = sr

    def predict(self, wave, sr):
        w = np.asarray(wave, dtype=np.float64)
        spec = np.abs(np.fft.rfft(w)) ** 2
        f = np.fft.rfftfreq(w.shape[0], 1.0 / self.sr)
        lo = spec[(f >= 100) & (f < 1500)].sum() + 1e-9
        hi = spec[(f >= 1500)].sum() + 1e-9
        mid = spec[(f >= 400) & (f < 900)].sum() + 1e-9
        z = 0.8 * np.log10(hi / lo) + 0.5 * np.log10(mid / lo) + 1.0 + 3.0 * float(np.mean(w[: w.shape[0] // 3] ** 2)) ** 0.5
        return float(1.0 / (1.0 + np.exp(-z)))


def test_track(sr: int, seconds: float, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    n = int(sr * seconds)
    t = np.arange(n) / sr
    y = 0.3 * np.sin(2 * np.pi * 220 * t) * (1 + 0.5 * np.sin(2 * np.pi * 0.7 * t))
    y += 0.2 * np.sin(2 * np.pi * (900 + 300 * t / seconds) * t)
    y += 0.1 * rng.standard_normal(n) * (t > seconds / 2)
    y += 0.05 * np.sin(2 * np.pi * 3100 * t) * (t < seconds / 3)
    return (0.5 * y / np.abs(y).max()).astype(np.float32)


def install_stubs(tracks: dict, written: list):
    sys.path.insert(0, str(REPO))
    from oracle import dsp

    librosa = types.ModuleType("librosa")
    librosa.load = lambda path, sr=None, duration=None, mono=True: (tracks[str(path)].copy(), sr)
    librosa.stft = lambda y, n_fft, hop_length, win_length, window="hann", center=True: dsp.stft(
        y, n_fft, hop_length, win_length).numpy()
    librosa.istft = lambda S, hop_length, win_length, window="hann", center=True: dsp.istft(
        S, hop_length, win_length).numpy()
    librosa.magphase = dsp.magphase
    librosa.fft_frequencies = lambda sr, n_fft: dsp.fft_frequencies(sr, n_fft)
    librosa.amplitude_to_db = lambda mag, ref=None: dsp.amplitude_to_db_refmax(mag)
    librosa.display = types.ModuleType("librosa.display")
    sf = types.ModuleType("soundfile")
    sf.write = lambda path, data, sr: written.append((Path(path).name, np.asarray(data).copy(), int(sr)))
    mods = {"librosa": librosa, "librosa.display": librosa.display, "soundfile": sf}
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.gridspec", "seaborn"):
        mods[name] = mock.MagicMock(name=name)
    sys.modules.update(mods)
    sys.path.insert(0, str(REF))


def main():
    sr = 16000
    tracks = {"/golden/trackA.wav": test_track(sr, 6.0, 1), "/golden/trackB.wav": test_track(sr, 4.1, 2)}
    written: list = []
    install_stubs(tracks, written)
    from src import spectrogram_explainability as ref_occ   # the reference's own module
    from src import dsp_band_ops as ref_fbp                   # the reference's own module

    GOLDEN.mkdir(parents=True, exist_ok=True)
    pred = EnergyPredictor(sr)

    # ---------------- occlusion: several (n_fft, hop, patch, stride) cases incl. half-even rounding ----
    occ_cases = [
        dict(track="/golden/trackA.wav", n_fft=512, hop=128, t_patch=128, t_stride=64, pf=20.0, sf=10.0, occ=0.0),
        dict(track="/golden/trackA.wav", n_fft=512, hop=128, t_patch=256, t_stride=256, pf=5.0, sf=5.0, occ=0.0),
        dict(track="/golden/trackB.wav", n_fft=256, hop=64, t_patch=300, t_stride=150, pf=33.0, sf=12.5, occ=0.0),
        dict(track="/golden/trackB.wav", n_fft=2048, hop=512, t_patch=1024, t_stride=1024, pf=20.0, sf=10.0, occ=0.0),
    ]
    out = {}
    for ci, c in enumerate(occ_cases):
        written.clear()
        ex = ref_occ.SpectrogramExplainability(
            predictor=pred, sr=sr, duration=120, n_fft=c["n_fft"], hop_length=c["hop"], win_length=c["n_fft"],
            spec_type="stft", top_n_windows=3, method="occlusion", use_original_audio=False,
            patch_time_frames=c["t_patch"], stride_time_frames=c["t_stride"],
            patch_freq_percent=c["pf"], stride_freq_percent=c["sf"])
        res = ex._compute_occlusion_map(c["track"], occlusion_value=c["occ"], baseline_threshold=0.0, verbose=False)
        pi = res.patch_importances
        out[f"occ{ci}_windows"] = np.array([[p["t_start"], p["t_end"], p["f_start"], p["f_end"]] for p in pi], np.int32)
        out[f"occ{ci}_importance"] = np.array([p["importance"] for p in pi], np.float64)
        out[f"occ{ci}_map"] = res.importance_map
        out[f"occ{ci}_baseline"] = np.float64(res.baseline_pred)
        with tempfile.TemporaryDirectory() as td:
            ex._save_top_occlusion_patches_from_list(y=res.y, S=res.S, patch_importances=pi, top_n=3,
                                                     save_dir=td, file_name="trk")
            groups = {}
            for g in ("all", "best", "worst", "most_influential"):
                with open(Path(td) / g / f"trk__{g}_occlusion_patches_from_list.json", encoding="utf-8") as f:
                    groups[g] = json.load(f)
        out[f"occ{ci}_groups_json"] = np.array(json.dumps(groups))
        out[f"occ{ci}_wav_names"] = np.array([w[0] for w in written])
        for wi, w in enumerate(written if ci in (0, 2) else []):   # window audio kept for two cases only
            out[f"occ{ci}_wav{wi}"] = w[1].astype(np.float32)
        out[f"occ{ci}_case_json"] = np.array(json.dumps(c))
    # ties: injected equal |importance| values must keep grid order in every group (stable sorts)
    tie_patches = [{"t_start": i, "t_end": i + 1, "f_start": 0, "f_end": 1, "importance": v}
                   for i, v in enumerate([0.5, -0.5, 0.25, 0.5, -0.25, 0.0, 0.25, -0.5, 0.0, 0.125])]
    ex = ref_occ.SpectrogramExplainability(predictor=pred, sr=sr, n_fft=512, hop_length=128, win_length=512,
                                           spec_type="stft", method="occlusion", use_original_audio=True)
    with tempfile.TemporaryDirectory() as td:
        ex._save_top_occlusion_patches_from_list(y=np.zeros(2048, np.float32), S=np.zeros((257, 17), np.complex64),
                                                 patch_importances=tie_patches, top_n=3, save_dir=td, file_name="tie")
        groups = {}
        for g in ("all", "best", "worst", "most_influential"):
            with open(Path(td) / g / f"tie__{g}_occlusion_patches_from_list.json", encoding="utf-8") as f:
                groups[g] = json.load(f)
    out["tie_importance"] = np.array([p["importance"] for p in tie_patches])
    out["tie_groups_json"] = np.array(json.dumps(groups))
    np.savez_compressed(GOLDEN / "ref_loops_occlusion.npz", **out)

    # ---------------- FBP ---------------------------------------------------------------------------
    out = {}
    fbp_cases = [
        dict(track="/golden/trackA.wav", preset="high_resolution", att=0.25, mode="rel", rel=0.2, mn=5.0, mx=500.0,
             hz=200.0, n_fft=2048, hop=512, norm=False, sr=16000),
        dict(track="/golden/trackB.wav", preset="default", att=0.0, mode="abs", rel=0.0, mn=0.0, mx=0.0,
             hz=200.0, n_fft=1024, hop=256, norm=True, sr=16000),
        dict(track="/golden/trackB.wav", preset="detailed_voice", att=0.5, mode="rel", rel=0.2, mn=5.0, mx=500.0,
             hz=0.0, n_fft=2048, hop=512, norm=True, sr=44100),
    ]
    for ci, c in enumerate(fbp_cases):
        fb = ref_fbp.FrequencyBandPerturbation(
            predictor=EnergyPredictor(c["sr"]), preset=c["preset"], attenuation=c["att"], transition_mode=c["mode"],
            transition_hz=c["hz"], transition_rel=c["rel"], transition_min_hz=c["mn"], transition_max_hz=c["mx"],
            sr=c["sr"], n_fft=c["n_fft"], hop_length=c["hop"], win_length=c["n_fft"], spec_type="stft",
            normalize_loudness=c["norm"])
        sig = tracks[c["track"]]
        res = fb._compute_component_importance(sig=sig, component_name="mixture", audio_path=c["track"])
        out[f"fbp{ci}_bands"] = np.array(fb.bands, np.float64)
        out[f"fbp{ci}_importance"] = np.array([b["importance"] for b in res.batch_importances])
        out[f"fbp{ci}_map"] = res.importance_map
        out[f"fbp{ci}_baseline"] = np.float64(res.baseline_pred)
        freqs = np.fft.rfftfreq(c["n_fft"], 1.0 / c["sr"])
        gains, trans_w = [], []
        for (lo, hi) in fb.bands:
            tr = fb._band_transition_width(lo, hi)
            keep = ref_fbp.smooth_band_keep_mask(freqs, lo, hi, trans=tr)
            gains.append(keep + c["att"] * (1.0 - keep))
            trans_w.append(tr)
        out[f"fbp{ci}_gain"] = np.array(gains)
        out[f"fbp{ci}_trans"] = np.array(trans_w)
        out[f"fbp{ci}_case_json"] = np.array(json.dumps(c))
    out["presets_json"] = np.array(json.dumps(ref_fbp.FREQUENCY_BAND_PRESETS))
    x = test_track(sr, 1.0, 7)
    out["match_rms_in"] = x
    out["match_rms_out"] = ref_fbp.match_rms(tracks["/golden/trackB.wav"], x * 0.3)
    np.savez_compressed(GOLDEN / "ref_loops_fbp.npz", **out)
    print("golden vectors written to", GOLDEN)


if __name__ == "__main__":
    main()
