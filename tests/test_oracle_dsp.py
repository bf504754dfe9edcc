"""Oracle DSP restatement cross-checked against torch.stft/istft and torchaudio (SURVEY 8c vi-viii)."""
import numpy as np
import torch

from audio_deepfake_explainability_b200.weights import ALPHA_120S, tiny_config
from oracle import dsp


def _y(n=40000, seed=0):
    rng = np.random.default_rng(seed)
    return (0.3 * rng.standard_normal(n)).astype(np.float32)


def test_stft_matches_torch_stft():
    y = _y()
    S = dsp.stft(y, 2048, 512, 2048)
    ref = torch.stft(torch.from_numpy(y), 2048, 512, 2048, window=torch.hann_window(2048, periodic=True),
                     center=True, pad_mode="constant", return_complex=True)
    assert S.shape == (1025, 1 + len(y) // 512)
    assert (S - ref).abs().max() < 2e-4 * ref.abs().max()


def test_istft_matches_torch_and_roundtrip():
    y = _y(512 * 60)
    S = dsp.stft(y, 2048, 512, 2048)
    yr = dsp.istft(S, 512, 2048)
    assert yr.shape[0] == 512 * (S.shape[1] - 1) == len(y)
    assert (yr - torch.from_numpy(y)).abs().max() < 2e-6
    ref = torch.istft(S, 2048, 512, 2048, window=torch.hann_window(2048, periodic=True), center=True)
    assert (yr - ref).abs().max() < 2e-6


def test_istft_linearity_and_patch_support():
    y = _y(512 * 80, 3)
    S = dsp.stft(y, 2048, 512, 2048)
    t0, t1, f0, f1 = 20, 40, 100, 151
    P = torch.zeros_like(S)
    P[f0:f1, t0:t1] = S[f0:f1, t0:t1]
    S_occ = S.clone()
    S_occ[f0:f1, t0:t1] = 0
    y_patch = dsp.istft(P, 512, 2048)
    assert (dsp.istft(S_occ, 512, 2048) - (dsp.istft(S, 512, 2048) - y_patch)).abs().max() < 1e-6
    nz = torch.nonzero(y_patch.abs() > 0).flatten()
    assert nz.min() >= t0 * 512 - 1024 and nz.max() < (t1 - 1) * 512 + 1024


def test_mel_frontend_restated_from_parts():
    cfg = ALPHA_120S
    y = torch.from_numpy(_y(16000 * 3, 5)).unsqueeze(0)
    power = dsp.mel_frontend(y, cfg, "power")
    # own restatement: reflect-pad STFT power x HTK filterbank
    spec = torch.stft(y[0], cfg.n_fft, cfg.hop_length, cfg.win_length, window=torch.hann_window(cfg.win_length),
                      center=True, pad_mode="reflect", return_complex=True).abs() ** 2
    mel = dsp.mel_filterbank(cfg).T @ spec
    assert torch.allclose(mel, power[0], rtol=1e-4, atol=1e-6)
    db = dsp.mel_frontend(y, cfg, "db")[0]
    ref_db = 10.0 * torch.log10(torch.clamp(power[0], min=1e-10))
    ref_db = torch.maximum(ref_db, ref_db.max() - 80.0)
    assert torch.allclose(db, ref_db, atol=1e-5)
    nrm = dsp.mel_frontend(y, cfg, "norm")[0]
    assert abs(float(nrm.mean())) < 1e-4 and abs(float(nrm.std()) - 1.0) < 1e-3


def test_batch_topdb_is_per_sample():
    cfg = tiny_config()
    a = torch.from_numpy(_y(8000, 1)).unsqueeze(0)
    b = 1e-3 * torch.from_numpy(_y(8000, 2)).unsqueeze(0)
    both = dsp.mel_frontend(torch.cat([a, b]), cfg, "db")
    assert torch.equal(both[1], dsp.mel_frontend(b, cfg, "db")[0])
