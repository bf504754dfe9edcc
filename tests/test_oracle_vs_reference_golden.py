"""The oracle restatement and the product's host-side integer logic, checked against golden vectors
produced by the reference's OWN loop code (oracle/make_golden_from_reference.py)."""
import json
import os

import numpy as np
import pytest

from audio_deepfake_explainability_b200 import grid
from oracle import dsp, loops
from oracle.make_golden_from_reference import EnergyPredictor, test_track as _track

G = os.path.join(os.path.dirname(__file__), "golden")
OCC = np.load(os.path.join(G, "ref_loops_occlusion.npz"))
FBP = np.load(os.path.join(G, "ref_loops_fbp.npz"))
SR = 16000
TRACKS = {"/golden/trackA.wav": _track(SR, 6.0, 1), "/golden/trackB.wav": _track(SR, 4.1, 2)}


@pytest.mark.parametrize("ci", [0, 1, 2, 3])
def test_occlusion_loop_matches_reference(ci):
    c = json.loads(str(OCC[f"occ{ci}_case_json"]))
    y = TRACKS[c["track"]]
    out = loops.occlusion_map(y, EnergyPredictor(SR), SR, c["n_fft"], c["hop"], c["n_fft"], c["t_patch"],
                              c["t_stride"], c["pf"], c["sf"], c["occ"], baseline_threshold=0.0)
    win = np.array([[p["t_start"], p["t_end"], p["f_start"], p["f_end"]] for p in out.patch_importances], np.int32)
    assert np.array_equal(win, OCC[f"occ{ci}_windows"])                     # indexing bit-exact
    imp = np.array([p["importance"] for p in out.patch_importances])
    assert np.array_equal(imp, OCC[f"occ{ci}_importance"])                  # same code path -> identical
    assert np.array_equal(out.importance_map, OCC[f"occ{ci}_map"])
    assert out.baseline_pred == float(OCC[f"occ{ci}_baseline"])
    # product host logic: the window grid
    n_freq, n_time = grid.stft_shape(len(y), c["n_fft"], c["hop"])
    assert np.array_equal(grid.occlusion_windows(n_freq, n_time, c["t_patch"], c["t_stride"], c["pf"], c["sf"]),
                          OCC[f"occ{ci}_windows"])
    assert np.array_equal(loops.saliency_from_windows(win, imp, n_freq, n_time), OCC[f"occ{ci}_map"])
    # top-k groups: oracle JSON payload and product index groups
    ref_groups = json.loads(str(OCC[f"occ{ci}_groups_json"]))
    got = loops.top_window_groups(out.patch_importances, 3, "trk", c["hop"], SR)
    assert got == ref_groups
    idx = grid.topk_window_groups(imp, 3)
    for g in ("all", "best", "worst", "most_influential"):
        ref_keys = [(w["t_start"], w["f_start"]) for w in ref_groups[g]["windows"]]
        assert [(int(win[i][0]), int(win[i][2])) for i in idx[g]] == ref_keys


@pytest.mark.parametrize("ci", [0, 2])
def test_top_window_audio_matches_reference(ci):
    c = json.loads(str(OCC[f"occ{ci}_case_json"]))
    y = TRACKS[c["track"]]
    S = dsp.stft(y, c["n_fft"], c["hop"], c["n_fft"]).numpy()
    groups = json.loads(str(OCC[f"occ{ci}_groups_json"]))
    wi = 0
    for g in ("best", "worst", "most_influential"):
        for w in groups[g]["windows"]:
            a = loops.window_audio(y, S, w, c["hop"], c["n_fft"], use_original_audio=False)
            assert np.array_equal(a.astype(np.float32), OCC[f"occ{ci}_wav{wi}"])
            name = str(OCC[f"occ{ci}_wav_names"][wi])
            assert name == (f"trk__{g}{w['rank']}_patch_{w['type']}_{w['abs_importance']:.3f}_"
                            f"t{w['t_start']}-{w['t_end']}_f{w['f_start']}-{w['f_end']}.wav")
            wi += 1


def test_stable_ties_match_reference():
    imp = OCC["tie_importance"]
    ref_groups = json.loads(str(OCC["tie_groups_json"]))
    idx = grid.topk_window_groups(imp, 3)
    for g in ("all", "best", "worst", "most_influential"):
        assert [int(i) for i in idx[g]] == [w["t_start"] for w in ref_groups[g]["windows"]], g
    patches = [{"t_start": i, "t_end": i + 1, "f_start": 0, "f_end": 1, "importance": float(v)} for i, v in enumerate(imp)]
    assert loops.top_window_groups(patches, 3, "tie", 128, SR) == ref_groups


@pytest.mark.parametrize("ci", [0, 1, 2])
def test_fbp_loop_matches_reference(ci):
    c = json.loads(str(FBP[f"fbp{ci}_case_json"]))
    bands = [tuple(b) for b in FBP[f"fbp{ci}_bands"]]
    sig = TRACKS[c["track"]]
    out = loops.fbp_component(sig, EnergyPredictor(c["sr"]), c["sr"], bands, c["att"], c["mode"], c["hz"], c["rel"],
                              c["mn"], c["mx"], c["n_fft"], c["hop"], c["n_fft"], c["norm"])
    assert np.array_equal(np.array([b["importance"] for b in out.batch_importances]), FBP[f"fbp{ci}_importance"])
    assert np.array_equal(out.importance_map, FBP[f"fbp{ci}_map"])
    assert out.baseline_pred == float(FBP[f"fbp{ci}_baseline"])
    # product host logic: gain table, transition widths, band->bin rows
    gain = grid.band_gain_table(bands, c["sr"], c["n_fft"], c["att"], c["mode"], c["rel"], c["mn"], c["mx"], c["hz"])
    assert np.array_equal(gain, FBP[f"fbp{ci}_gain"])
    tw = [grid.band_transition_width(lo, hi, c["mode"], c["rel"], c["mn"], c["mx"], c["hz"]) for lo, hi in bands]
    assert np.array_equal(np.array(tw), FBP[f"fbp{ci}_trans"])
    rows = grid.band_bin_ranges(bands, c["sr"], c["n_fft"])
    ref_map = FBP[f"fbp{ci}_map"]
    rebuilt = np.zeros_like(ref_map)
    for (b0, b1), d in zip(rows, FBP[f"fbp{ci}_importance"]):
        rebuilt[b0:b1, :] += d
    assert np.array_equal(rebuilt, ref_map)


def test_presets_and_match_rms():
    ref = json.loads(str(FBP["presets_json"]))
    assert {k: [list(b) for b in v] for k, v in grid.FREQUENCY_BAND_PRESETS.items()} == ref
    got = dsp.match_rms(TRACKS["/golden/trackB.wav"], FBP["match_rms_in"] * 0.3)
    assert np.array_equal(got, FBP["match_rms_out"])
