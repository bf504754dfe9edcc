"""LayerNorm / head / saliency / band-map / ranking kernels vs plain references (through the C ABI)."""
import numpy as np
import pytest
import torch

from audio_deepfake_explainability_b200 import grid
from gpu_util import D, P, lib, ok
from oracle import loops

pytestmark = pytest.mark.gpu


def test_layernorm_bf16_and_inplace_groups():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2 * 1376 + 3, 384, generator=g) * 3 + 1
    ga, ba = torch.randn(384, generator=g), torch.randn(384, generator=g)
    gb, bb = torch.randn(384, generator=g), torch.randn(384, generator=g)
    out = torch.zeros(x.shape, dtype=torch.bfloat16, device="cuda")
    ok(lib().b200x_layernorm(P(D(x)), x.shape[0], 384, P(D(ga)), P(D(ba)), P(None), P(None), 0, 0, 1e-5, P(out), P(None), 0, P(None)))
    ref = torch.nn.functional.layer_norm(x, (384,), ga, ba, 1e-5)
    assert (out.float().cpu() - ref).abs().max().item() < 8e-3 * ref.abs().max().item()   # bf16 output: 2^-8 relative
    xi = x[: 2 * 1376].cuda().clone()
    ok(lib().b200x_layernorm(P(xi), 2 * 1376, 384, P(D(ga)), P(D(ba)), P(D(gb)), P(D(bb)), 1376, 1248, 1e-6, P(None), P(xi), 0, P(None)))
    xr = x[: 2 * 1376].reshape(2, 1376, 384)
    ref2 = torch.cat([torch.nn.functional.layer_norm(xr[:, :1248], (384,), ga, ba, 1e-6),
                      torch.nn.functional.layer_norm(xr[:, 1248:], (384,), gb, bb, 1e-6)], dim=1).reshape(-1, 384)
    assert (xi.cpu() - ref2).abs().max().item() < 2e-5


@pytest.mark.parametrize("use_norm", [1, 0])
def test_head_matches_reference(use_norm):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(3, 1376, 384, generator=g)
    ga, ba, w = torch.randn(384, generator=g), torch.randn(384, generator=g), torch.randn(384, generator=g) * 0.05
    part = torch.zeros(3 * lib().b200x_head_slices(), device="cuda")
    logit, prob = torch.zeros(3, device="cuda"), torch.zeros(3, device="cuda")
    ok(lib().b200x_head(P(D(x)), 3, 1376, 384, P(D(ga)), P(D(ba)), 1e-5, use_norm, P(D(w)), 0.3, P(part), P(logit), P(prob), P(None)))
    f = torch.nn.functional.layer_norm(x, (384,), ga, ba, 1e-5) if use_norm else x
    ref = f.mean(1) @ w + 0.3
    assert (logit.cpu() - ref).abs().max().item() < 1e-5
    assert (prob.cpu() - torch.sigmoid(ref)).abs().max().item() < 1e-6


@pytest.mark.parametrize("stride_t,sf", [(1024, 5.0), (512, 2.5), (256, 1.25)])
def test_saliency_map_bit_exact(stride_t, sf):
    n_freq, n_time = 1025, 3751
    wins = grid.occlusion_windows(n_freq, n_time, 1024, stride_t, 5.0, sf)
    rng = np.random.default_rng(len(wins))
    delta = rng.standard_normal(len(wins)) * 1e-2
    out = torch.full((n_freq, n_time), float("nan"), dtype=torch.float64, device="cuda")
    ok(lib().b200x_saliency_reduce(P(D(wins)), P(D(delta)), len(wins), n_freq, n_time, P(out), P(None)))
    ref = loops.saliency_from_windows(wins, delta, n_freq, n_time)
    assert np.array_equal(out.cpu().numpy(), ref)                            # float64, same order -> identical bits


def test_saliency_map_empty_and_ragged():
    out = torch.full((7, 130), float("nan"), dtype=torch.float64, device="cuda")
    ok(lib().b200x_saliency_reduce(P(None), P(None), 0, 7, 130, P(out), P(None)))
    assert (out == 0).all()
    wins = np.array([[0, 130, 0, 7], [129, 130, 6, 7], [3, 3, 1, 2]], np.int32)
    d = np.array([0.5, -0.25, 9.0])
    ok(lib().b200x_saliency_reduce(P(D(wins)), P(D(d)), 3, 7, 130, P(out), P(None)))
    assert np.array_equal(out.cpu().numpy(), loops.saliency_from_windows(wins, d, 7, 130))


def test_band_map_bit_exact():
    bands = grid.FREQUENCY_BAND_PRESETS["high_resolution"]
    rows = grid.band_bin_ranges(bands, 16000, 2048)
    delta = np.random.default_rng(3).standard_normal(len(bands))
    out = torch.zeros((1025, 500), dtype=torch.float64, device="cuda")
    ok(lib().b200x_band_map(P(D(rows)), P(D(delta)), len(bands), 1025, 500, P(out), P(None)))
    freqs = grid.fft_frequencies(16000, 2048)
    ref = np.zeros((1025, 500))
    for (lo, hi), d in zip(bands, delta):
        ref[(freqs >= lo) & (freqs <= hi), :] += d
    assert np.array_equal(out.cpu().numpy(), ref)


@pytest.mark.parametrize("mode,key,desc", [(0, abs, True), (1, abs, False), (2, float, True), (3, float, False)])
def test_rank_is_stable_like_python_sorted(mode, key, desc):
    rng = np.random.default_rng(mode)
    v = np.round(rng.standard_normal(825), 1)                                # many exact ties
    v[::50] = 0.0
    order = torch.zeros(len(v), dtype=torch.int32, device="cuda")
    ok(lib().b200x_rank(P(D(v)), len(v), mode, P(order), P(None)))
    ref = sorted(range(len(v)), key=lambda i: key(v[i]), reverse=desc)
    assert order.cpu().tolist() == ref


def test_delta_is_float64_difference():
    p = torch.tensor([0.25, 0.5000001, 0.9], device="cuda")
    d = torch.zeros(3, dtype=torch.float64, device="cuda")
    ok(lib().b200x_delta(P(p), 0.5, 3, P(d), P(None)))
    ref = np.float64(np.float32(0.5)) - p.cpu().numpy().astype(np.float64)
    assert np.array_equal(d.cpu().numpy(), ref)
