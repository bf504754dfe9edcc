#!/usr/bin/env python
"""Benchmark of the perturbation-explainability hot path (BASELINE.json: occluded-spectrogram evals/sec).

    python bench.py --gpus N --steps K --warmup W             # this engine (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...    # the reference's CPU path (oracle port), rank 0 only
    python bench.py --scaling strong ...                       # headline = BASELINE configs[3], windows sharded over the ranks
    python bench.py --workload fbp64 | stems1000 ...           # headline = BASELINE configs[2] / configs[4]

Headline (default): a "step" is one pass of the hot path over one synthetic 120 s / 16 kHz track (BASELINE configs[1]):
baseline prediction + the 228-window occlusion sweep (1024-frame x 5 % patches, half-window stride) + the delta-prob ->
saliency reduction + the four stable top-k rankings + patch-only iSTFT reconstruction of the top windows.  With N GPUs every
rank sweeps its own track (weak scaling, no data-path collective: the tracks are unrelated).

The same JSON line also carries, measured in the same run (short, bounded):
  strong    : BASELINE configs[3] - dense quarter-stride sweep (825 windows) of one track per generator family (5 tracks),
              each track's windows sharded over the N ranks, ONE NCCL all-gather of the per-shard probabilities per track
              on a side stream (overlapping the next track's sweep), then delta -> saliency -> rankings on every rank
  workloads : configs[2] (FBP, 13-band high_resolution bank, 64 tracks) and configs[4] (1000 AudioLIME stem masks)
  roofline  : dominant kernel (attention) against the measured bf16 peak, and `hbm_stages`: STFT / masked iSTFT / mel
              kernels against the measured HBM bandwidth (dense and sparse launches)
  topk      : k-th-boundary margins of the four top-k groups and their agreement with the cached CPU-oracle deltas

  value : evals/s with the track, its STFT and the window list already resident in HBM (device-side timing, CUDA
          events on the engine stream, max over ranks)
  e2e   : the same metric through the public API with HOST buffers: wave upload, window upload, probabilities /
          saliency map / top-window audio read back, inside the timed region.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SR = 16000
DURATION = 120.0
PATCH_T, STRIDE_T, PATCH_F, STRIDE_F = 1024, 512, 5.0, 2.5             # configs[1]: half-window stride, 228 windows
DENSE_STRIDE_T, DENSE_STRIDE_F = 256, 1.25                             # configs[3]: quarter-window stride, 825 windows
TOP_N = 5
TIE_EPS = 1e-4
# algorithmic work per perturbed evaluation (2 flops per MAC, padding excluded)
FLOP_ATTN_PER_EVAL = 12 * 2 * (2 * 1376 * 1376 * 64 * 6)            # QK^T + PV, 12 layers
FLOP_GEMM_PER_EVAL = 12 * 2 * 1376 * (384 * 1152 + 384 * 384 + 2 * 384 * 1025) + 2 * (1248 * 384 * 384 + 128 * 3744 * 384)
FLOP_PER_EVAL = FLOP_ATTN_PER_EVAL + FLOP_GEMM_PER_EVAL               # 81.1 GFLOP
METRIC = "occluded-spectrogram evals/sec"


def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p.get("hbm_gbs", 6650.0), "bf16_tflops": p.get("bf16_tflops", 1590.0),
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", 1400.0), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self._proc, self._t = index, [], None, None

    def _run(self):
        for line in self._proc.stdout:                       # one line per 20 ms sample from ONE long-running nvidia-smi
            cells = [c.strip() for c in line.strip().split(",")]
            if len(cells) >= 6:
                self.rows.append(cells)

    def __enter__(self):
        try:
            self._proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index),
                                           "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
            t_end = time.time() + 8.0                        # nvidia-smi needs up to a few seconds to start on an 8-GPU box
            while not self.rows and time.time() < t_end:
                time.sleep(0.02)
            self.rows.clear()                                # samples from here on fall inside the timed region
        except Exception:
            self._proc = None
        return self

    def __exit__(self, *a):
        if self._proc is not None:
            self._proc.terminate()
            try:
                self._proc.wait(timeout=5)
            except Exception:
                self._proc.kill()
            self._t.join(timeout=5)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


# --------------------------------------------------------------------------------------------- workload descriptions
def workload_config(args):
    """`config` of the JSON line: depends on the command line only, so both arms (--impl engine / reference) print the same."""
    n = args.gpus
    if args.workload == "fbp64":
        return {"workload": "configs[2]: FBP sweep, 13-band high_resolution bank, attenuation 0.25, 64 synthetic 120 s 16 kHz tracks "
                            "(13 band copies + 1 baseline per track = 896 evals), random-init SpecTTTra-alpha-120s",
                "tracks": 64, "bands": 13, "copies_per_chunk": 224, "parallelism": f"tracks sharded over {n} GPU(s), no collective",
                "l2_policy": "inputs larger than L2 (224-copy activation set, ~2.2 GB per layer)"}
    if args.workload == "stems1000":
        return {"workload": "configs[4]: AudioLIME-style perturbation, 1000 stem-mask recombinations of 4 precomputed synthetic stems "
                            "of one 120 s 16 kHz track (np.random.RandomState(0) bits, row 0 all ones), every row evaluated",
                "masks": 1000, "stems": 4, "copies_per_chunk": args.chunk,
                "parallelism": f"mask rows sharded over {n} GPU(s), one all-gather of probabilities",
                "l2_policy": "inputs larger than L2 (chunk activation set ~2.3 GB per layer)"}
    if args.scaling == "strong":
        return {"workload": "configs[3]: dense occlusion sweep (1024-frame x 5% window, quarter-window stride: 825 windows) of one "
                            "synthetic 120 s 16 kHz track per generator family (REAL/SUNO/SUNO_PRO/UDIO/ElevenLabs), "
                            "+ baseline + saliency map + rankings per track",
                "windows_per_track": 825, "tracks": 5, "copies_per_chunk": args.chunk,
                "parallelism": f"each track's windows sharded over {n} GPU(s); one NCCL all-gather of per-shard probabilities per track "
                               "on a side stream, overlapping the next track's sweep",
                "l2_policy": "inputs larger than L2 (chunk activation set ~1-2.3 GB per layer)"}
    return {"workload": "configs[1]: occlusion sweep, one synthetic 120 s 16 kHz track per GPU, random-init SpecTTTra-alpha-120s, "
                        "1024-frame x 5% window, half-window stride (228 evals) + baseline + saliency map + top-5 window iSTFT",
            "windows_per_track": 228, "tracks": n, "copies_per_chunk": args.chunk,
            "l2_policy": "inputs larger than L2: every step streams the 229-copy activation set (~2.3 GB: residual stream, "
                         "qkv, attention, MLP hidden) plus 228 x 1035 spectrogram rows through the 126 MB L2",
            "parallelism": f"one track per GPU over {n} GPU(s) (independent tracks: no data-path collective)"}


def fbp_tracks(n_tracks: int, rank: int = 0, world: int = 1):
    """64 distinct synthetic tracks for configs[2]: 8 synthesised tracks x 8 circular time shifts (synthesis costs ~1 s of
    host time per track; a shift changes every STFT frame).  Returns this rank's tracks (sharded by track)."""
    from audio_deepfake_explainability_b200 import grid, synth
    base = [synth.synth_track(synth.FAMILIES[i % 5], i // 5, SR, DURATION) for i in range(min(8, n_tracks))]
    lo, hi = grid.shard_range(n_tracks, rank, world)
    out = [np.roll(base[i % len(base)], (i // len(base)) * 117649) for i in range(lo, hi)]
    return np.stack(out) if out else np.zeros((0, len(base[0])), np.float32)


# --------------------------------------------------------------------------------------------- CPU baseline / reference arm
def oracle_evals_per_s(n_evals: int, threads: int, workload: str = "occlusion", stride=(STRIDE_T, STRIDE_F)):
    """Times the reference's CPU path (oracle port: librosa-style iSTFT + SpecTTTra forward, fp32, batch 1) on a bounded
    sample of the same workload: ``n_evals`` perturbed evaluations, one warm-up evaluation discarded."""
    import torch
    from audio_deepfake_explainability_b200 import grid, synth
    from audio_deepfake_explainability_b200.weights import ALPHA_120S, random_state_dict
    from oracle import dsp, spectttra                      # CPU baseline leg: the one place bench.py executes oracle/

    torch.set_num_threads(threads)
    sd = random_state_dict(ALPHA_120S, 0)
    pred = spectttra.OraclePredictor(sd, ALPHA_120S, "fp32")
    if workload == "stems1000":
        y, stems = synth.synth_track("REAL", 0, SR, DURATION, with_stems=True)
        st = np.stack([stems[k] for k in sorted(stems)])
        masks = np.random.RandomState(0).randint(0, 2, 1000 * 4).reshape(1000, 4)
        masks[0] = 1

        def one(i):
            return pred.predict((masks[i % 1000][:, None] * st).sum(0).astype(np.float32), SR)
    elif workload == "fbp64":
        y = synth.synth_track("REAL", 0, SR, DURATION)
        S = dsp.stft(y).numpy()
        gains = grid.band_gain_table(grid.FREQUENCY_BAND_PRESETS["high_resolution"], SR, 2048, 0.25, "rel", 0.2, 5.0, 500.0, 0.0)

        def one(i):
            return pred.predict(dsp.istft(S * gains[i % 13][:, None]).numpy(), SR)
    else:
        y = synth.synth_track("REAL", 0, SR, DURATION)
        S = dsp.stft(y).numpy()
        wins = grid.occlusion_windows(S.shape[0], S.shape[1], PATCH_T, stride[0], PATCH_F, stride[1])

        def one(i):
            t0, t1, f0, f1 = wins[i % len(wins)]
            patch = S[f0:f1, t0:t1].copy()
            S[f0:f1, t0:t1] = 0.0
            y_occ = dsp.istft(S).numpy()
            S[f0:f1, t0:t1] = patch
            return pred.predict(y_occ, SR)

    one(0)
    t = time.perf_counter()
    for i in range(n_evals):
        one(1 + i)
    dt = time.perf_counter() - t
    return n_evals / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    threads = os.cpu_count() or 1
    per_step = 4                                            # bounded sample: 4 perturbed evals per step
    stride = (DENSE_STRIDE_T, DENSE_STRIDE_F) if args.scaling == "strong" else (STRIDE_T, STRIDE_F)
    for _ in range(args.warmup):
        oracle_evals_per_s(1, threads, args.workload, stride)
    times = []
    for _ in range(args.steps):
        _, dt = oracle_evals_per_s(per_step, threads, args.workload, stride)
        times.append(dt)
    value = per_step * len(times) / sum(times)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{per_step} perturbed evals per step (oracle port: librosa-style iSTFT + SpecTTTra forward fp32, batch 1; "
                                   "torch.istft-based, i.e. faster than the reference's numpy iSTFT) of the workload in `config`; "
                                   "parity of the port's DSP / classifier modules is unpinned (librosa / sonics absent)"},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# --------------------------------------------------------------------------------------------- engine arm
class Ctx:
    """Process-wide handles shared by the workload functions."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from audio_deepfake_explainability_b200 import _lib
        from audio_deepfake_explainability_b200.engine import Engine
        from audio_deepfake_explainability_b200.weights import ALPHA_120S, random_state_dict

        self.torch, self.dist, self._lib = torch, dist, _lib
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise RuntimeError("bench.py needs a CUDA device: the engine has no CPU fallback")
        torch.cuda.set_device(self.local)
        if self.world > 1:
            if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
                os.environ["NCCL_DEBUG"] = "WARN"          # keep stdout to the one JSON line the driver parses
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        self.args = args
        self.cfg = ALPHA_120S
        self.sd = random_state_dict(self.cfg, 0)
        self.eng = Engine(self.cfg, self.sd, copies_per_chunk=args.chunk, max_samples=int(SR * DURATION), device=self.local)
        if args.no_alternate:
            self.eng.set_alternate(False)
        self.lib = self.eng.lib
        self.stream = torch.cuda.ExternalStream(self.eng.stream, device=torch.device("cuda", self.local))
        self.side = torch.cuda.Stream(device=torch.device("cuda", self.local))
        self.sp = C.c_void_p(self.eng.stream)

    def P(self, t):
        return C.c_void_p(t.data_ptr())

    def check(self, status, what):
        self._lib.check(status, what)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()
        self.eng.synchronize()

    def timed(self, fn, steps):
        """steps x fn() bracketed by barrier + synchronize; device time from CUDA events on the engine stream, max over ranks."""
        torch = self.torch
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(self.stream)
        out = None
        for _ in range(steps):
            out = fn()
        self.side.synchronize()
        e1.record(self.stream)
        self.eng.synchronize()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        ms = max(e0.elapsed_time(e1), 0.0)
        ms, wall_ms = self.max_over_ranks(ms), self.max_over_ranks(1e3 * wall)
        self.barrier()
        return ms, wall_ms, out

    def max_over_ranks(self, v: float) -> float:
        if self.world == 1:
            return v
        t = self.torch.tensor([v], device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def close(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()
        self.eng.close()


def occlusion_weak(ctx: Ctx, steps: int, warmup: int):
    """configs[1] on every rank's own track."""
    torch, eng, lib = ctx.torch, ctx.eng, ctx.lib
    from audio_deepfake_explainability_b200 import grid, synth
    P, check, sp = ctx.P, ctx.check, ctx.sp

    fam = synth.FAMILIES[ctx.rank % len(synth.FAMILIES)]
    y = synth.synth_track(fam, ctx.rank // len(synth.FAMILIES), SR, DURATION)
    n_freq, n_time = grid.stft_shape(len(y), 2048, 512)
    windows = grid.occlusion_windows(n_freq, n_time, PATCH_T, STRIDE_T, PATCH_F, STRIDE_F)
    n_win = len(windows)
    assert n_win == 228
    y_pin = torch.from_numpy(y).pin_memory()
    win_pin = torch.from_numpy(windows).pin_memory()
    d_wave = torch.from_numpy(y).cuda()
    d_win = torch.from_numpy(windows).cuda()
    d_prob = torch.zeros(n_win, device="cuda")
    d_base = torch.zeros(1, device="cuda")
    d_delta = torch.zeros(n_win, dtype=torch.float64, device="cuda")
    d_map = torch.zeros(n_freq, n_time, dtype=torch.float64, device="cuda")
    d_order = torch.zeros(4, n_win, dtype=torch.int32, device="cuda")
    h_order = torch.zeros(4, n_win, dtype=torch.int32).pin_memory()
    torch.cuda.synchronize()
    state = {}

    def device_step():
        check(lib.b200x_engine_set_track(eng._h, P(d_wave), len(y), 1), "set_track")
        # baseline prediction + the 228 occluded copies in ONE device pass (the track rides as copy 229 of the chunk)
        check(lib.b200x_engine_occlusion_sweep_base(eng._h, P(d_win), n_win, 0.0, 1, P(d_prob), P(d_base)), "sweep")
        check(lib.b200x_delta_dev(P(d_prob), P(d_base), n_win, P(d_delta), sp), "delta")
        check(lib.b200x_saliency_reduce(P(d_win), P(d_delta), n_win, n_freq, n_time, P(d_map), sp), "saliency")
        for mode in range(4):
            check(lib.b200x_rank(P(d_delta), n_win, mode, P(d_order[mode]), sp), "rank")
        with torch.cuda.stream(ctx.stream):                 # the copy is ordered behind the rank kernels on the engine stream
            h_order.copy_(d_order, non_blocking=True)
        eng.synchronize()
        order = h_order.numpy()
        top = np.unique(np.concatenate([order[0][:TOP_N], order[1][:TOP_N], order[2][:TOP_N], order[3][:TOP_N]]))
        eng.window_audio(windows[top])
        return 6                                            # delta + saliency + 4 rank kernels launched outside the engine's own counter

    def host_step():
        eng.set_track(y_pin.numpy())
        prob, base = eng.occlusion_sweep(win_pin.numpy(), 0.0, with_baseline=True)
        base = float(base)
        delta = np.float64(np.float32(base)) - prob.astype(np.float64)
        sal = eng.saliency_map(windows, delta)
        orders = [eng.rank(delta, m) for m in range(4)]
        top = np.unique(np.concatenate([o[:TOP_N] for o in orders]))
        aud = eng.window_audio(windows[top])
        state["delta"], state["base"] = delta, base
        h2d = y.nbytes + windows.nbytes + windows.nbytes + delta.nbytes + 4 * delta.nbytes + windows[top].nbytes
        d2h = prob.nbytes + 4 + sal.nbytes + 4 * 4 * n_win + sum(a.nbytes for a in aud)
        return h2d, d2h

    for _ in range(max(warmup, 3)):
        device_step()
    l0 = eng.launch_count
    with ClockSampler(ctx.local) as clk:
        ms_dev, _, extra = ctx.timed(device_step, steps)
    launches = (eng.launch_count - l0) + extra * steps
    for _ in range(2):
        host_step()
    ms_host, wall_host, io = ctx.timed(host_step, steps)
    ms_host = max(ms_host, wall_host)                       # the e2e number is host-visible time: the calls block on the host
    evals = (n_win + 1) * ctx.world * steps
    res = {"value": evals / (ms_dev * 1e-3), "ms_per_step": ms_dev / steps, "evals_per_step": (n_win + 1) * ctx.world,
           "e2e": {"value": evals / (ms_host * 1e-3), "unit": "evals/s", "h2d_bytes_per_step": int(io[0]), "d2h_bytes_per_step": int(io[1]),
                   "ms_per_step": ms_host / steps},
           "launches": int(launches), "clocks": clk.summary(), "family": fam, "delta": state.get("delta"), "base": state.get("base"),
           "device_step": device_step, "n_win": n_win}
    return res


def occlusion_strong(ctx: Ctx, steps: int, warmup: int):
    """configs[3]: 5 tracks x 825 windows; every track's windows are sharded over the ranks (grid.shard_range of the t-major
    window list, src/spectrogram_explainability.py:644-648); one all-gather of the per-shard probabilities per track."""
    torch, dist, eng, lib = ctx.torch, ctx.dist, ctx.eng, ctx.lib
    from audio_deepfake_explainability_b200 import grid, synth
    from audio_deepfake_explainability_b200.sonics_api import B200Predictor
    from audio_deepfake_explainability_b200.spectrogram_explainability import SpectrogramExplainability
    P, check = ctx.P, ctx.check
    rank, world = ctx.rank, ctx.world

    tracks = [synth.synth_track(f, 0, SR, DURATION) for f in synth.FAMILIES]
    n_freq, n_time = grid.stft_shape(len(tracks[0]), 2048, 512)
    windows = grid.occlusion_windows(n_freq, n_time, PATCH_T, DENSE_STRIDE_T, PATCH_F, DENSE_STRIDE_F)
    n_win = len(windows)
    assert n_win == 825
    lo, hi = grid.shard_range(n_win, rank, world)
    n_loc = hi - lo
    slot = (n_win + world - 1) // world + 1                 # padded shard + one slot for the baseline (filled by rank 0)
    d_waves = [torch.from_numpy(t).cuda() for t in tracks]
    d_win_all = torch.from_numpy(windows).cuda()
    d_win_loc = d_win_all[lo:hi].contiguous()
    nt = len(tracks)
    d_send = [torch.zeros(slot, device="cuda") for _ in range(nt)]
    d_recv = [torch.zeros(world * slot, device="cuda") for _ in range(nt)]
    d_prob = [torch.zeros(n_win, device="cuda") for _ in range(nt)]
    d_delta = [torch.zeros(n_win, dtype=torch.float64, device="cuda") for _ in range(nt)]
    d_map = torch.zeros(n_freq, n_time, dtype=torch.float64, device="cuda")
    d_order = [torch.zeros(4, n_win, dtype=torch.int32, device="cuda") for _ in range(nt)]
    index = torch.cat([torch.arange(r * slot, r * slot + (grid.shard_range(n_win, r, world)[1] - grid.shard_range(n_win, r, world)[0]))
                       for r in range(world)]).cuda()
    side_p = C.c_void_p(ctx.side.cuda_stream)
    done = [torch.cuda.Event() for _ in range(nt)]
    torch.cuda.synchronize()

    def device_step():
        for f in range(nt):
            check(lib.b200x_engine_set_track(eng._h, P(d_waves[f]), len(tracks[f]), 1), "set_track")
            if rank == 0:                                   # rank 0's shard carries the unperturbed track as one more copy
                check(lib.b200x_engine_occlusion_sweep_base(eng._h, P(d_win_loc), n_loc, 0.0, 1, P(d_send[f]), P(d_send[f][slot - 1:])), "sweep")
            elif n_loc:
                check(lib.b200x_engine_occlusion_sweep(eng._h, P(d_win_loc), n_loc, 0.0, 1, P(d_send[f])), "sweep")
            done[f].record(ctx.stream)
            with torch.cuda.stream(ctx.side):               # gather + reductions of track f overlap the sweep of track f + 1
                ctx.side.wait_event(done[f])
                if world > 1:
                    dist.all_gather_into_tensor(d_recv[f], d_send[f])
                else:
                    d_recv[f].copy_(d_send[f])
                torch.index_select(d_recv[f], 0, index, out=d_prob[f])
                check(lib.b200x_delta_dev(P(d_prob[f]), P(d_recv[f][slot - 1:]), n_win, P(d_delta[f]), side_p), "delta")
                check(lib.b200x_saliency_reduce(P(d_win_all), P(d_delta[f]), n_win, n_freq, n_time, P(d_map), side_p), "saliency")
                for mode in range(4):
                    check(lib.b200x_rank(P(d_delta[f]), n_win, mode, P(d_order[f][mode]), side_p), "rank")
        ctx.side.synchronize()
        return 6 * nt

    pred = object.__new__(B200Predictor)                    # the public explainer API around the bench's own engine
    pred.engine, pred.cfg, pred.model_name, pred.device = eng, ctx.cfg, "random-init-seed0", "cuda"
    ex = SpectrogramExplainability(pred, sr=SR, duration=int(DURATION), spec_type="stft", method="occlusion", use_original_audio=False,
                                   patch_time_frames=PATCH_T, stride_time_frames=DENSE_STRIDE_T, patch_freq_percent=PATCH_F,
                                   stride_freq_percent=DENSE_STRIDE_F, top_n_windows=TOP_N)

    def host_step():
        h2d = d2h = 0
        for y in tracks:
            r = ex.occlusion_map_from_wave(y, baseline_threshold=0.0 if world == 1 else 1e-9, verbose=False, want_spectrogram=False)
            ex.top_window_groups(r.patch_importances, TOP_N, "track")
            h2d += y.nbytes + 2 * windows.nbytes + 5 * 8 * n_win
            d2h += 4 * n_win + 4 + r.importance_map.nbytes + 16 * n_win
        return h2d, d2h

    for _ in range(max(1, warmup)):
        device_step()
    l0 = eng.launch_count
    ms_dev, _, extra = ctx.timed(device_step, steps)
    launches = (eng.launch_count - l0) + extra * steps
    host_step()
    ms_host, wall_host, io = ctx.timed(host_step, steps)
    ms_host = max(ms_host, wall_host)
    evals = nt * (n_win + 1) * steps
    return {"value": evals / (ms_dev * 1e-3), "ms_per_step": ms_dev / steps, "evals_per_step": nt * (n_win + 1),
            "e2e": {"value": evals / (ms_host * 1e-3), "unit": "evals/s", "h2d_bytes_per_step": int(io[0]), "d2h_bytes_per_step": int(io[1]),
                    "ms_per_step": ms_host / steps},
            "launches": int(launches), "windows_per_rank": n_loc}


def fbp64(ctx: Ctx, steps: int, warmup: int, n_tracks: int = 64):
    """configs[2]: the 13-band high_resolution bank on 64 tracks through the batch-of-tracks entry point (host buffers in,
    probabilities out: this IS the end-to-end path; the device-resident number is the same call minus the wave upload)."""
    from audio_deepfake_explainability_b200 import grid
    from audio_deepfake_explainability_b200.engine import Engine
    eng = Engine(ctx.cfg, ctx.sd, copies_per_chunk=224, max_samples=int(SR * DURATION), device=ctx.local)
    try:
        waves = fbp_tracks(n_tracks, ctx.rank, ctx.world)
        bands = grid.FREQUENCY_BAND_PRESETS["high_resolution"]
        gains = grid.band_gain_table(bands, SR, 2048, 0.25, "rel", 0.2, 5.0, 500.0, 0.0).astype(np.float32)
        rows = grid.band_bin_ranges(bands, SR, 2048)
        out = {}
        n_freq, n_time = 1025, 1 + waves.shape[1] // 512
        for normalize, materialize in ((False, False), (True, False), (False, True)):
            def step():
                base, prob = eng.fbp_sweep_tracks(waves, gains, normalize)
                delta = base.astype(np.float64)[:, None] - prob.astype(np.float64)
                m = None
                for d in delta:
                    # the batch API's default: a read-only broadcast view of the map's one distinct column (host, band-order
                    # float64 adds); materialize = the device-built [1025, n_time] float64 array read back per track
                    m = eng.band_map(rows, d) if materialize else grid.band_map_view(rows, d, n_freq, n_time)
                return m
            for _ in range(max(2, warmup)):
                step()
            l0 = eng.launch_count
            ctx.barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                step()
            eng.synchronize()
            wall = ctx.max_over_ranks(1e3 * (time.perf_counter() - t0))
            evals = n_tracks * (len(bands) + 1) * steps
            key = "materialized_maps" if materialize else f"normalize_loudness={normalize}"
            out[key] = {"value": evals / (wall * 1e-3), "unit": "evals/s", "ms_per_step": wall / steps,
                        "evals_per_step": n_tracks * (len(bands) + 1),
                        "gpu_launches": int(eng.launch_count - l0),
                        "h2d_bytes_per_step": int(waves.nbytes + gains.nbytes),
                        "d2h_bytes_per_step": int(len(waves) * (14 * 4 + (n_freq * n_time * 8 if materialize else 0)))}
        return out
    finally:
        eng.close()


def stems1000(ctx: Ctx, steps: int, warmup: int, n_masks: int = 1000):
    """configs[4]: 1000 stem-mask recombinations of one track's four stems; every row is evaluated (the LIME explainer itself
    de-duplicates to the <= 16 distinct masks - reported separately)."""
    from audio_deepfake_explainability_b200 import dist as xdist, grid, synth
    eng = ctx.eng
    y, stems = synth.synth_track("REAL", 0, SR, DURATION, with_stems=True)
    st = np.stack([stems[k] for k in sorted(stems)])
    masks = np.random.RandomState(0).randint(0, 2, n_masks * 4).reshape(n_masks, 4).astype(np.uint8)
    masks[0] = 1
    lo, hi = grid.shard_range(n_masks, ctx.rank, ctx.world)

    def step():
        local = eng.stem_sweep(st, masks[lo:hi]) if hi > lo else np.zeros(0, np.float32)
        return xdist.gather_shards(local, n_masks)

    for _ in range(max(1, warmup)):
        step()
    l0 = eng.launch_count
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        p = step()
    eng.synchronize()
    wall = ctx.max_over_ranks(1e3 * (time.perf_counter() - t0))
    uniq = np.unique(masks, axis=0)
    t1 = time.perf_counter()
    eng.stem_sweep(st, uniq)
    dedup_ms = 1e3 * (time.perf_counter() - t1)
    return {"value": n_masks * steps / (wall * 1e-3), "unit": "evals/s", "ms_per_step": wall / steps, "evals_per_step": n_masks,
            "gpu_launches": int(eng.launch_count - l0), "h2d_bytes_per_step": int(st.nbytes + masks[lo:hi].nbytes),
            "d2h_bytes_per_step": int(4 * (hi - lo)), "distinct_masks": int(len(uniq)), "dedup_sweep_ms": dedup_ms,
            "mean_prob": float(np.mean(p))}


def hbm_stages(ctx: Ctx, peaks, copies: int = 64):
    """STFT / masked iSTFT / mel kernels alone (kernel-level C ABI, CUDA events, `copies` copies per launch): algorithmic bytes
     / time against the measured HBM bandwidth.  dense = every frame of every copy; sparse = the occlusion path
    (only the classifier frames [t0 - 4, t1 + 4) an occlusion window can change: 1032 of 3751)."""
    torch, lib, P = ctx.torch, ctx.lib, ctx.P
    L = int(SR * DURATION)
    n_frames = 1 + L // 512
    g = torch.Generator(device="cuda").manual_seed(0)
    wave = torch.randn(L, device="cuda", generator=g) * 0.1
    S = torch.zeros(n_frames, 1028, 2, device="cuda")
    y = torch.zeros(copies, L + 8, device="cuda")
    wins = torch.tensor([[1024, 2048, 100, 151]] * copies, dtype=torch.int32, device="cuda")
    n_cta = -(-n_frames // lib.b200x_mel_frames_per_cta())
    db = torch.zeros(copies, n_frames, 128, device="cuda")
    cmax = torch.zeros(copies, n_cta, device="cuda")
    rng = torch.zeros(copies, 2, dtype=torch.int32, device="cuda")
    img_t = torch.zeros(copies, 3744, 128, dtype=torch.bfloat16, device="cuda")
    part = torch.zeros(copies * 64, dtype=torch.float64, device="cuda")
    fl = torch.zeros(copies, device="cuda")
    null = C.c_void_p(0)
    ctx.check(lib.b200x_stft(P(wave), L, 2048, 512, 0, P(S), 1028, null), "stft")
    ctx.check(lib.b200x_frame_ranges(P(wins), copies, n_frames, P(rng), null), "frame_ranges")
    sparse_frames = 1032

    def run(fn, iters=5):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / iters              # us

    k = {}
    k["stft"] = (run(lambda: ctx.check(lib.b200x_stft(P(wave), L, 2048, 512, 0, P(S), 1028, null), "stft"), 20), 4 * L + 8 * 1025 * n_frames)
    k["istft_masked_dense"] = (run(lambda: ctx.check(lib.b200x_istft_masked(P(S), 1028, n_frames, copies, 1, P(wins), 0.0, null, P(y), L + 8, null,
                                                                            null, 0, null), "istft")), copies * (8 * 1025 * n_frames + 4 * L))
    k["mel_db_dense"] = (run(lambda: ctx.check(lib.b200x_mel_db(P(y), L + 8, L, copies, 16000, 128, 20.0, 8000.0, 1e-10, null, 0.0, L, P(db),
                                                                n_frames, P(cmax), null, 0, null), "mel")), copies * (4 * L + 4 * 128 * n_frames))
    k["normalize_resize"] = (run(lambda: ctx.check(lib.b200x_mel_normalize_resize(P(db), n_frames, P(cmax), n_cta, copies, n_frames, 128, 80.0, 1,
                                                                                  1e-6, 3744, null, null, null, null, P(part), P(fl), P(img_t),
                                                                                  null, 3744, null), "resize")),
                             copies * (4 * 128 * n_frames + 2 * 128 * 3744))          # one image: the engine's 128-mel path
    k["istft_masked_sparse"] = (run(lambda: ctx.check(lib.b200x_istft_masked(P(S), 1028, n_frames, copies, 1, P(wins), 0.0, null, P(y), L + 8, null,
                                                                             P(rng), sparse_frames, null), "istft")),
                                copies * (8 * 1025 * (sparse_frames + 3) + 4 * 512 * (sparse_frames + 3)))
    k["mel_db_sparse"] = (run(lambda: ctx.check(lib.b200x_mel_db(P(y), L + 8, L, copies, 16000, 128, 20.0, 8000.0, 1e-10, null, 0.0, L, P(db),
                                                                 n_frames, P(cmax), P(rng), sparse_frames, null), "mel")),
                          copies * (4 * 512 * (sparse_frames + 3) + 4 * 128 * sparse_frames))
    out = {"copies_per_launch": copies, "peak_gbs": peaks["hbm_gbs"], "peak_source": f"{peaks['source']} (copy bandwidth)",
           "note": "algorithmic bytes per launch (SURVEY 8d) / CUDA-event time; the FFT-based kernels are bound by fp32 issue, not by HBM: "
                   "a 2048-point real FFT is ~65 kFLOP per frame = 25 FLOP per algorithmic byte, i.e. 163 TFLOP/s would be needed at the "
                   "HBM roofline against a 72 TFLOP/s fp32 peak (DESIGN.md 3.3)"}
    # the bound that applies to the FFT kernels: fp32 issue.  65 kFLOP per 2048-point real frame (5 N log2 N of the 1024-point
    # complex FFT + real-FFT pre / post pass + window + power / mel or overlap-add) against 148 SMs x 128 lanes x 2 x 1.965 GHz
    FRAME_FLOP, FP32_PEAK = 65.0e3, 148 * 128 * 2 * 1.965e9 / 1e12
    frames = {"stft": n_frames, "istft_masked_dense": copies * n_frames, "mel_db_dense": copies * n_frames,
              "istft_masked_sparse": copies * (sparse_frames + 3), "mel_db_sparse": copies * sparse_frames}
    out["fp32_peak_tflops"] = FP32_PEAK
    for name, (us, nbytes) in k.items():
        gbs = nbytes / us / 1e3
        out[name] = {"us": us, "algorithmic_bytes": int(nbytes), "gbs": gbs, "frac": gbs / peaks["hbm_gbs"]}
        if name in frames:
            tf = frames[name] * FRAME_FLOP / us / 1e6
            out[name].update(bound="fp32", frames=int(frames[name]), fp32_tflops=tf, fp32_frac=tf / FP32_PEAK)
        else:
            out[name]["bound"] = "hbm"
    return out


def topk_report(delta, family: str):
    """k-th-boundary margins of the engine's own deltas and agreement with the cached full-size oracle deltas
    (tests/golden/fullsize_occlusion_<family>0.npz, made by oracle/make_golden_fullsize.py; a fixture read, no oracle code runs)."""
    from audio_deepfake_explainability_b200 import grid
    if delta is None:
        return None
    out = {"top_n": TOP_N, "margins_engine": grid.topk_boundary_margins(delta, TOP_N), "tie_epsilon": TIE_EPS}
    path = os.path.join(ROOT, "tests", "golden", f"fullsize_occlusion_{family}0.npz")
    if os.path.exists(path):
        z = np.load(path)
        for mode in ("fp32", "bf16"):
            if f"delta_{mode}" not in z.files:
                continue
            ref = z[f"delta_{mode}"]
            err = np.abs(delta - ref)
            big = np.abs(ref) > 1e-3
            ge, gr = grid.topk_window_groups(delta, TOP_N), grid.topk_window_groups(ref, TOP_N)
            se, sr_ = grid.topk_window_groups(delta, TOP_N, TIE_EPS), grid.topk_window_groups(ref, TOP_N, TIE_EPS)
            names = ("best", "worst", "most_influential")
            raw = {g: {"set_equal": set(ge[g].tolist()) == set(gr[g].tolist()), "order_equal": bool(np.array_equal(ge[g], gr[g]))} for g in names}
            snap = {g: {"set_equal": set(se[g].tolist()) == set(sr_[g].tolist()), "order_equal": bool(np.array_equal(se[g], sr_[g]))} for g in names}
            out[f"vs_oracle_{mode}"] = {"max_abs_err": float(err.max()), "max_rel_err_where_abs_gt_1e-3": float((err[big] / np.abs(ref[big])).max()) if big.any() else None,
                                        "margins_oracle": grid.topk_boundary_margins(ref, TOP_N), "groups_raw": raw,
                                        "groups_tie_epsilon": snap}
    return out


def run_engine(args):
    ctx = Ctx(args)
    peaks = read_peaks()
    eng, world, rank = ctx.eng, ctx.world, ctx.rank
    extras = not args.no_extras
    line = {"metric": METRIC, "unit": "evals/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(args)}
    roofline = None
    if args.workload == "fbp64":
        r = fbp64(ctx, args.steps, args.warmup)
        head = r["normalize_loudness=False"]
        line.update(value=head["value"], ms_per_step=head["ms_per_step"], gpu_launches=head["gpu_launches"],
                    e2e={"value": head["value"], "unit": "evals/s", "h2d_bytes_per_step": head["h2d_bytes_per_step"],
                         "d2h_bytes_per_step": head["d2h_bytes_per_step"]}, detail=r)
    elif args.workload == "stems1000":
        r = stems1000(ctx, args.steps, args.warmup)
        line.update(value=r["value"], ms_per_step=r["ms_per_step"], gpu_launches=r["gpu_launches"],
                    e2e={"value": r["value"], "unit": "evals/s", "h2d_bytes_per_step": r["h2d_bytes_per_step"],
                         "d2h_bytes_per_step": r["d2h_bytes_per_step"]}, detail=r)
    else:
        weak = None
        if args.scaling == "strong":
            r = occlusion_strong(ctx, args.steps, args.warmup)
            line.update(value=r["value"], ms_per_step=r["ms_per_step"], gpu_launches=r["launches"], e2e=r["e2e"],
                        windows_per_rank=r["windows_per_rank"])
            with ClockSampler(ctx.local) as clk:
                occlusion_strong(ctx, 1, 0)
            line["clocks"] = clk.summary()
        else:
            weak = occlusion_weak(ctx, args.steps, args.warmup)
            line.update(value=weak["value"], ms_per_step=weak["ms_per_step"], gpu_launches=weak["launches"], e2e=weak["e2e"],
                        clocks=weak["clocks"])
            if extras:
                s = occlusion_strong(ctx, max(1, min(args.steps, 2)), 1)
                line["strong"] = {"workload": "configs[3]: 5 tracks x 825 windows (quarter stride), each track's windows sharded over the "
                                              "ranks, one NCCL all-gather per track on a side stream; + baseline, saliency map, rankings",
                                  "value": s["value"], "unit": "evals/s", "ms_per_step": s["ms_per_step"], "evals_per_step": s["evals_per_step"],
                                  "windows_per_rank": s["windows_per_rank"], "e2e": s["e2e"], "gpu_launches": s["launches"], "scaling": "strong"}

        # ---- roofline of the dominant kernel: per-class CUDA-event timing over one more timed pass of configs[1] --------
        if weak is None:
            weak = occlusion_weak(ctx, 1, 3)
        n_win = weak["n_win"]
        eng.set_timing(True)
        n_pass = max(1, min(args.steps, 3))
        for _ in range(n_pass):
            weak["device_step"]()
        tim = eng.get_timing()
        eng.set_timing(False)
        total_ms = sum(v[0] for v in tim.values())
        shares = {k: (v[0] / total_ms if total_ms else 0.0) for k, v in tim.items()}
        dom = "attention"                                   # one launch per layer; the "gemm" class is four different problems
        evals_timed = (n_win + 1) * n_pass
        flops = FLOP_ATTN_PER_EVAL * evals_timed
        dom_ms, dom_n = tim[dom]
        achieved = flops / (dom_ms * 1e-3) / 1e12 if dom_ms else 0.0
        gemm_ms = tim["gemm"][0]
        copies_per_launch = (n_win + 1) * 12.0 * n_pass / dom_n if dom_n else float(n_win + 1)
        traffic, traffic_src = None, None
        for tname in ("r02_ncu_traffic.json", "r01_p_ncu_traffic.json"):
            tpath = os.path.join(ROOT, "profiles", tname)
            if os.path.exists(tpath):
                with open(tpath) as f:
                    tj = json.load(f)
                t = tj.get("attention_kernel")
                if t:
                    traffic = t["traffic_bytes_per_launch"] * copies_per_launch / float(tj.get("_captured_copies_per_launch", n_win + 1))
                    traffic_src = f"profiles/{tname} (dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full of bench.py)"
                    break
        ev_s = weak["value"] / world
        roofline = {"bound": "tensor", "kernel": "attention (fused QK^T / softmax / PV, one launch per layer)",
                    "achieved": achieved, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                    "frac": achieved / peaks["bf16_tflops_sustained"], "traffic": traffic, "traffic_source": traffic_src,
                    "algorithmic_bytes_per_launch": copies_per_launch * 1376 * (1152 + 384) * 2, "copies_per_launch": copies_per_launch,
                    "peak_source": f"{peaks['source']} (sustained bf16; kernel timed inside a long step)",
                    "avg_launch_ms": dom_ms / dom_n if dom_n else None, "launches": dom_n,
                    "algorithmic_flops_per_launch": flops / dom_n if dom_n else None,
                    "gemm_class": {"achieved": FLOP_GEMM_PER_EVAL * evals_timed / (gemm_ms * 1e-3) / 1e12 if gemm_ms else None,
                                   "unit": "TFLOP/s", "launches": tim["gemm"][1],
                                   "frac": (FLOP_GEMM_PER_EVAL * evals_timed / (gemm_ms * 1e-3) / 1e12 / peaks["bf16_tflops_sustained"]) if gemm_ms else None},
                    "share_of_step": shares, "ms_per_class": {k: v[0] / n_pass for k, v in tim.items()},
                    "whole_forward_frac_of_peak": FLOP_PER_EVAL * ev_s / 1e12 / peaks["bf16_tflops_sustained"]}
        if extras and world == 1:
            roofline["hbm_stages"] = hbm_stages(ctx, peaks)
        if rank == 0:
            line["topk"] = topk_report(weak.get("delta"), weak["family"])
        if extras:
            f = fbp64(ctx, 1, 2)
            s = stems1000(ctx, 1, 1)
            line["workloads"] = {"fbp64": dict(f["normalize_loudness=False"], workload="configs[2]: 13-band high_resolution FBP, 64 tracks, "
                                               "batch-of-tracks entry point, host buffers (end to end); importance maps as read-only "
                                               "broadcast views of their one distinct column (the batch API's default)",
                                               normalize_loudness_true=f["normalize_loudness=True"],
                                               materialized_maps=f["materialized_maps"]),
                                 "stems1000": dict(s, workload="configs[4]: 1000 stem-mask recombinations of 4 stems, host buffers (end to end)")}
    if roofline is not None:
        line["roofline"] = roofline
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            import torch as _t
            threads = os.cpu_count() or 1
            stride = (DENSE_STRIDE_T, DENSE_STRIDE_F) if args.scaling == "strong" else (STRIDE_T, STRIDE_F)
            cpu_rate, cpu_dt = oracle_evals_per_s(args.cpu_evals, threads, args.workload, stride)
            cpu = {"value": cpu_rate, "unit": "evals/s", "cores": _t.get_num_threads(), "kind": "port",
                   "sample": f"{args.cpu_evals} perturbed evals of the workload in `config` (oracle port: iSTFT + SpecTTTra fp32, batch 1), {cpu_dt:.1f} s; "
                             "parity of the port's DSP / classifier modules is unpinned (librosa / sonics absent from the image)"}
        line["cpu_baseline"] = cpu
        emit(line)
    # every torch tensor that was used on the engine's stream must be released BEFORE the engine destroys that stream
    # (the caching allocator records an event on each stream a block was used on when the block is freed)
    weak = None
    import gc
    gc.collect()
    ctx.torch.cuda.synchronize()
    ctx.torch.cuda.empty_cache()
    ctx.close()


_RESULT_OUT = None


def emit(line: dict) -> None:
    """The ONE JSON line of the contract, on the process's real stdout."""
    out = _RESULT_OUT if _RESULT_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    # stdout carries exactly one JSON line: everything libraries print there (NCCL's version banner under torchrun, ...)
    # is sent to stderr by pointing fd 1 at fd 2 and keeping a private duplicate of the real stdout for emit()
    global _RESULT_OUT
    sys.stdout.flush()
    _RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: configs[1], one track per rank (headline); strong: configs[3], windows of each track sharded over the ranks")
    ap.add_argument("--workload", default="occlusion", choices=["occlusion", "fbp64", "stems1000"])
    ap.add_argument("--chunk", type=int, default=229, help="copies per pass (one chunk = the whole 228-window sweep + the unperturbed track)")
    ap.add_argument("--cpu-evals", type=int, default=10, help="bounded CPU-baseline sample (perturbed evals)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the strong / workloads / hbm_stages blocks (headline only)")
    ap.add_argument("--no-alternate", action="store_true", help="diagnostic: every kernel walks its rows / tiles forward")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_engine(args)
    _RESULT_OUT.flush()
    sys.stderr.flush()
    os._exit(0)       # the line is out and every rank has left its last barrier: skip interpreter teardown (NCCL / CUDA atexit order)


if __name__ == "__main__":
    main()
