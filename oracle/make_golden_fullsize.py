"""ORACLE tooling (test infrastructure only): full-size golden vectors of the CPU oracle.

    python -m oracle.make_golden_fullsize [occlusion] [fbp] [shipped]

The CPU oracle needs 0.3-0.6 s per perturbed evaluation at BASELINE.json's sizes, so the full sweeps cannot run inside
the GPU test session.  This script runs them once in the build container and caches the per-window / per-band
delta-probabilities under tests/golden/; the `-m gpu` tests compare the engine against ALL of them
(tests/test_gpu_fullsize_golden.py) and the top-k group orderings derived from them.

  occlusion : configs[1] sweep (1024 x 5 %, half stride: 228 windows) on the bench track (REAL 0) and the test track
              (SUNO 0), with the fp32 oracle (the reference's arithmetic) and the bf16-GEMM-input oracle (the engine's
              arithmetic contract)
  fbp       : configs[2]: the 13-band high_resolution bank, attenuation 0.25, normalize_loudness False and True (SUNO 0)
  shipped   : the reference's own shipped configuration (spectrogram_explainability.yaml:35-63): sr 44 100, 120 s,
              n_time 10 336, 1024 x 20 % windows at 10 % stride -> 90 windows; 12 of them + the baseline

Same arithmetic as `oracle/loops.py::occlusion_map` / `fbp_component` (one evaluation per window, batch 1, restore
after each window: src/spectrogram_explainability.py:663-703, src/dsp_band_ops.py:573-653).
"""
from __future__ import annotations

import sys
import time
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parents[1]
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))
GOLDEN = REPO / "tests" / "golden"

from audio_deepfake_explainability_b200 import grid, synth                               # noqa: E402
from audio_deepfake_explainability_b200.weights import ALPHA_120S, random_state_dict    # noqa: E402
from oracle import dsp, loops, spectttra                                                # noqa: E402

CFG = ALPHA_120S
SHIPPED_PICK = (0, 7, 13, 22, 31, 38, 44, 45, 52, 66, 77, 89)


def occlusion_deltas(y, sr, wins, pred, pick=None):
    S = dsp.stft(y).numpy()
    base = pred.predict(y, sr)
    out = []
    idx = range(len(wins)) if pick is None else pick
    for n, i in enumerate(idx):
        t0, t1, f0, f1 = (int(v) for v in wins[i])
        patch = S[f0:f1, t0:t1].copy()
        S[f0:f1, t0:t1] = 0.0
        y_occ = dsp.istft(S).numpy()
        S[f0:f1, t0:t1] = patch
        # trim / zero-pad to len(y) (src/spectrogram_explainability.py:676-680)
        if len(y_occ) > len(y):
            y_occ = y_occ[: len(y)]
        elif len(y_occ) < len(y):
            y_occ = np.pad(y_occ, (0, len(y) - len(y_occ)))
        out.append(base - pred.predict(y_occ, sr))
        if n % 20 == 0:
            print(f"    window {n + 1}/{len(idx)}", flush=True)
    return float(base), np.asarray(out, np.float64)


def make_occlusion(sd):
    for fam in ("REAL", "SUNO"):
        y = synth.synth_track(fam, 0, 16000, 120.0)
        n_freq, n_time = grid.stft_shape(len(y), 2048, 512)
        wins = grid.occlusion_windows(n_freq, n_time, 1024, 512, 5.0, 2.5)
        res = {"windows": wins}
        for mode in ("fp32", "bf16"):
            t = time.time()
            base, d = occlusion_deltas(y, 16000, wins, spectttra.OraclePredictor(sd, CFG, mode))
            res[f"base_{mode}"] = np.float64(base)
            res[f"delta_{mode}"] = d
            print(f"  {fam} {mode}: base {base:.6f}, max |delta| {np.abs(d).max():.3e}, {time.time() - t:.0f} s", flush=True)
        np.savez_compressed(GOLDEN / f"fullsize_occlusion_{fam}0.npz", **res)


def make_fbp(sd):
    from audio_deepfake_explainability_b200.dsp_band_ops import FREQUENCY_BAND_PRESETS
    y = synth.synth_track("SUNO", 0, 16000, 120.0)
    bands = FREQUENCY_BAND_PRESETS["high_resolution"]
    pred = spectttra.OraclePredictor(sd, CFG, "fp32")
    res = {"bands": np.asarray(bands, np.int64)}
    for normalize in (False, True):
        t = time.time()
        r = loops.fbp_component(y, pred, 16000, bands=bands, attenuation=0.25, transition_mode="rel", transition_rel=0.2,
                                transition_min_hz=5.0, transition_max_hz=500.0, normalize_loudness=normalize)
        res[f"delta_norm{int(normalize)}"] = np.asarray([b["importance"] for b in r.batch_importances], np.float64)
        res[f"base_norm{int(normalize)}"] = np.float64(r.baseline_pred)
        print(f"  fbp normalize={normalize}: {time.time() - t:.0f} s", flush=True)
    np.savez_compressed(GOLDEN / "fullsize_fbp_SUNO0.npz", **res)


def make_shipped(sd):
    sr = 44100
    y = synth.synth_track("UDIO", 0, sr, 120.0)
    n_freq, n_time = grid.stft_shape(len(y), 2048, 512)
    wins = grid.occlusion_windows(n_freq, n_time, 1024, 1024, 20.0, 10.0)
    assert (n_time, len(wins)) == (10336, 90), (n_time, len(wins))
    pred = spectttra.OraclePredictor(sd, CFG, "fp32")
    base, d = occlusion_deltas(y, sr, wins, pred, SHIPPED_PICK)
    np.savez_compressed(GOLDEN / "shipped_config_44k_UDIO0.npz", windows=wins, pick=np.asarray(SHIPPED_PICK), base_fp32=np.float64(base),
                        delta_fp32=d)
    print(f"  shipped: base {base:.6f}, deltas {d}", flush=True)


def main():
    torch.set_num_threads(max(1, (__import__("os").cpu_count() or 2)))
    sd = random_state_dict(CFG, 0)
    what = set(sys.argv[1:]) or {"occlusion", "fbp", "shipped"}
    GOLDEN.mkdir(parents=True, exist_ok=True)
    if "shipped" in what:
        make_shipped(sd)
    if "fbp" in what:
        make_fbp(sd)
    if "occlusion" in what:
        make_occlusion(sd)


if __name__ == "__main__":
    main()
