#!/usr/bin/env python
"""Wall-clock (synchronised) time of each phase of bench.py's device-resident step: python tools/phase_times.py [chunk]"""
import ctypes as C, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audio_deepfake_explainability_b200 import _lib, grid, synth
from audio_deepfake_explainability_b200.engine import Engine
from audio_deepfake_explainability_b200.weights import ALPHA_120S, random_state_dict

chunk = int(sys.argv[1]) if len(sys.argv) > 1 else 16
SR, DUR = 16000, 120.0
eng = Engine(ALPHA_120S, random_state_dict(ALPHA_120S, 0), copies_per_chunk=chunk, max_samples=int(SR * DUR), device=0)
lib = eng.lib
y = synth.synth_track("REAL", 0, SR, DUR)
n_freq, n_time = grid.stft_shape(len(y), 2048, 512)
windows = grid.occlusion_windows(n_freq, n_time, 1024, 512, 5.0, 2.5)
n_win = len(windows)
d_wave = torch.from_numpy(y).cuda(); d_win = torch.from_numpy(windows).cuda()
d_prob = torch.zeros(n_win, device="cuda"); d_base = torch.zeros(1, device="cuda")
d_delta = torch.zeros(n_win, dtype=torch.float64, device="cuda")
d_map = torch.zeros(n_freq, n_time, dtype=torch.float64, device="cuda")
d_order = torch.zeros(4, n_win, dtype=torch.int32, device="cuda")
P = lambda t: C.c_void_p(t.data_ptr()); sp = C.c_void_p(eng.stream)

def sync():
    eng.synchronize(); torch.cuda.synchronize()

phases = {}
def timed(name, fn):
    sync(); t = time.perf_counter(); fn(); sync(); phases.setdefault(name, []).append((time.perf_counter() - t) * 1e3)

for it in range(4):
    timed("set_track", lambda: _lib.check(lib.b200x_engine_set_track(eng._h, P(d_wave), len(y), 1), "x"))
    timed("predict(1)", lambda: _lib.check(lib.b200x_engine_predict(eng._h, P(d_wave), len(y), 1, 1, P(d_base), None), "x"))
    timed("sweep(228)", lambda: _lib.check(lib.b200x_engine_occlusion_sweep(eng._h, P(d_win), n_win, 0.0, 1, P(d_prob)), "x"))
    def red():
        base = float(d_base.item())
        _lib.check(lib.b200x_delta(P(d_prob), base, n_win, P(d_delta), sp), "x")
        _lib.check(lib.b200x_saliency_reduce(P(d_win), P(d_delta), n_win, n_freq, n_time, P(d_map), sp), "x")
        for mode in range(4):
            _lib.check(lib.b200x_rank(P(d_delta), n_win, mode, P(d_order[mode]), sp), "x")
    timed("delta+saliency+rank", red)
    order = d_order.cpu().numpy()
    top = np.unique(np.concatenate([order[m][:5] for m in range(4)]))
    timed(f"window_audio({len(top)})", lambda: eng.window_audio(windows[top]))
    timed("sweep(16 only)", lambda: _lib.check(lib.b200x_engine_occlusion_sweep(eng._h, P(d_win), 16, 0.0, 1, P(d_prob)), "x"))
    timed("sweep(224)", lambda: _lib.check(lib.b200x_engine_occlusion_sweep(eng._h, P(d_win), 224, 0.0, 1, P(d_prob)), "x"))
for k, v in phases.items():
    print(f"{k:24s} " + " ".join(f"{x:8.2f}" for x in v) + " ms")
eng.set_timing(True)
_lib.check(lib.b200x_engine_occlusion_sweep(eng._h, P(d_win), 224, 0.0, 1, P(d_prob)), "x")
tim = eng.get_timing(); eng.set_timing(False)
print("per-class kernel ms for sweep(224):", {k: round(v[0], 2) for k, v in tim.items()}, "sum", round(sum(v[0] for v in tim.values()), 2))
eng.close()
