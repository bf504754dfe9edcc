#!/usr/bin/env python
"""FBP throughput (BASELINE configs[2]: high_resolution bank, attenuation 0.25, N synthetic 120 s tracks on one B200):
python tools/fbp_bench.py [tracks] [normalize 0/1]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audio_deepfake_explainability_b200 import grid, synth
from audio_deepfake_explainability_b200.engine import Engine
from audio_deepfake_explainability_b200.weights import ALPHA_120S, random_state_dict

n_tracks = int(sys.argv[1]) if len(sys.argv) > 1 else 8
normalize = bool(int(sys.argv[2])) if len(sys.argv) > 2 else False
SR = 16000
tracks = [synth.synth_track(synth.FAMILIES[i % 5], i // 5, SR, 120.0) for i in range(min(n_tracks, 5))]
eng = Engine(ALPHA_120S, random_state_dict(ALPHA_120S, 0), copies_per_chunk=16, max_samples=len(tracks[0]))
gains = grid.band_gain_table(grid.FREQUENCY_BAND_PRESETS["high_resolution"], SR, 2048, 0.25, "rel", 0.2, 5.0, 500.0, 0.0).astype(np.float32)
rows = grid.band_bin_ranges(grid.FREQUENCY_BAND_PRESETS["high_resolution"], SR, 2048)


def one(y):
    eng.set_track(y)
    base = float(eng.predict_track())
    prob = eng.fbp_sweep(gains, normalize)
    delta = np.float64(np.float32(base)) - prob.astype(np.float64)
    return eng.band_map(rows, delta)


for y in tracks[:2]:
    one(y)
eng.synchronize()
t0 = time.perf_counter()
for i in range(n_tracks):
    one(tracks[i % len(tracks)])
eng.synchronize()
dt = time.perf_counter() - t0
print(f"{n_tracks} tracks x 14 evals (13 bands + baseline), normalize_loudness={normalize}: {1e3 * dt / n_tracks:.2f} ms per track, "
      f"{14 * n_tracks / dt:.0f} evals/s end to end (host buffers)")
eng.set_timing(True)
one(tracks[0])
print({k: round(v[0], 3) for k, v in eng.get_timing().items()})
eng.close()

# the same work through the batch-of-tracks entry point (band copies of ~17 tracks per launch)
eng = Engine(ALPHA_120S, random_state_dict(ALPHA_120S, 0), copies_per_chunk=224, max_samples=len(tracks[0]))
waves = np.stack([tracks[i % len(tracks)] for i in range(n_tracks)])


def batch():
    base, prob = eng.fbp_sweep_tracks(waves, gains, normalize)
    delta = base.astype(np.float64)[:, None] - prob.astype(np.float64)
    for d in delta:                        # maps are consumed one at a time (keeping 64 x 30 MB alive only measures page faults)
        m = eng.band_map(rows, d)
    return m


batch()
batch()                                    # second call captures the CUDA graphs of the new chunk shapes
eng.synchronize()
t0 = time.perf_counter()
base, prob = eng.fbp_sweep_tracks(waves, gains, normalize)
eng.synchronize()
t_sweep = time.perf_counter() - t0
print(f"batched sweep alone: {1e3 * t_sweep / n_tracks:.2f} ms per track, {14 * n_tracks / t_sweep:.0f} evals/s")
t0 = time.perf_counter()
batch()
eng.synchronize()
dt = time.perf_counter() - t0
print(f"batched: {n_tracks} tracks x 14 evals: {1e3 * dt / n_tracks:.2f} ms per track, {14 * n_tracks / dt:.0f} evals/s end to end (host buffers)")
