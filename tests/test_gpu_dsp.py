"""DSP kernels (STFT, masked iSTFT/OLA, mel front-end, normalise+resize) vs the oracle, through the C ABI."""
import numpy as np
import pytest
import torch

from audio_deepfake_explainability_b200 import grid, synth
from audio_deepfake_explainability_b200.weights import ALPHA_120S
from gpu_util import D, P, lib, ok
from oracle import dsp, spectttra

pytestmark = pytest.mark.gpu
STRIDE = 1028


def _track(seconds=8.0, family="REAL", extra=0):
    y = synth.synth_track(family, 0, 16000, seconds)
    return y[: len(y) - extra] if extra else y


def _gpu_stft(y, reflect=0):
    n_frames = 1 + len(y) // 512
    S = torch.zeros(n_frames, STRIDE, 2, device="cuda")
    ok(lib().b200x_stft(P(D(y)), len(y), 2048, 512, reflect, P(S), STRIDE, P(None)))
    return S


@pytest.mark.parametrize("extra", [0, 100])
def test_stft_matches_oracle(extra):
    y = _track(6.0, extra=extra)
    S = _gpu_stft(y)
    got = torch.view_as_complex(S[:, :1025].contiguous()).cpu().T           # [n_freq, n_frames]
    ref = dsp.stft(y)
    assert got.shape == ref.shape
    assert (got - ref).abs().max().item() < 2e-5 * ref.abs().max().item()


def _gpu_istft(S, n_frames, copies, mode, windows=None, occ=0.0, gains=None, sumsq=False):
    L = 512 * (n_frames - 1)
    y = torch.full((copies, L + 8), float("nan"), device="cuda")
    ss = torch.zeros(copies, dtype=torch.float64, device="cuda") if sumsq else None
    w = torch.from_numpy(np.ascontiguousarray(windows, dtype=np.int32)).cuda() if windows is not None else None
    g = torch.from_numpy(np.ascontiguousarray(gains, dtype=np.float32)).cuda() if gains is not None else None
    ok(lib().b200x_istft_masked(P(S), STRIDE, n_frames, copies, mode, P(w), occ, P(g), P(y), L + 8, P(ss), P(None), 0, P(None)))
    return y[:, :L].cpu(), (ss.cpu() if sumsq else None)


def test_istft_roundtrip_and_oracle():
    y = _track(7.3, "UDIO")
    y = y[: 512 * (len(y) // 512)]
    S = _gpu_stft(y)
    n_frames = S.shape[0]
    got, ss = _gpu_istft(S, n_frames, 1, 0, sumsq=True)
    assert (got[0] - torch.from_numpy(y)).abs().max().item() < 3e-6        # STFT/iSTFT round trip
    ref = dsp.istft(dsp.stft(y))
    assert (got[0] - ref).abs().max().item() < 3e-6
    assert abs(ss[0].item() - float((got[0].double() ** 2).sum())) < 1e-6 * float((got[0].double() ** 2).sum())


def test_istft_occlusion_windows_match_oracle():
    y = _track(10.0, "SUNO")
    S = _gpu_stft(y)
    n_frames = S.shape[0]
    wins = np.array([[0, 64, 0, 51], [100, 228, 500, 551], [250, n_frames, 974, 1025], [30, 31, 0, 1025], [5, 5, 3, 3]], np.int32)
    got, _ = _gpu_istft(S, n_frames, len(wins), 1, windows=wins, occ=0.0)
    S_ref = dsp.stft(y)
    for i, (t0, t1, f0, f1) in enumerate(wins):
        S_occ = S_ref.clone()
        S_occ[f0:f1, t0:t1] = 0.0
        ref = dsp.istft(S_occ)
        assert (got[i] - ref).abs().max().item() < 3e-6, i
    # non-zero occlusion value and the keep-only (top-window audio) mode
    got_v, _ = _gpu_istft(S, n_frames, 1, 1, windows=wins[1:2], occ=0.25)
    S_occ = S_ref.clone(); S_occ[500:551, 100:228] = 0.25
    assert (got_v[0] - dsp.istft(S_occ)).abs().max().item() < 3e-6
    got_k, _ = _gpu_istft(S, n_frames, 2, 3, windows=wins[:2])
    for i in range(2):
        t0, t1, f0, f1 = wins[i]
        Pm = torch.zeros_like(S_ref); Pm[f0:f1, t0:t1] = S_ref[f0:f1, t0:t1]
        assert (got_k[i] - dsp.istft(Pm)).abs().max().item() < 3e-6


def test_istft_band_gain_matches_oracle_fp64():
    y = _track(6.0, "ElevenLabs")
    S = _gpu_stft(y)
    n_frames = S.shape[0]
    bands = grid.FREQUENCY_BAND_PRESETS["high_resolution"]
    gains = grid.band_gain_table(bands, 16000, 2048, 0.25, "rel", 0.2, 5.0, 500.0, 200.0)
    got, _ = _gpu_istft(S, n_frames, len(bands), 2, gains=gains)
    S_ref = dsp.stft(y).numpy()
    mag, phase = dsp.magphase(S_ref)
    for i in range(len(bands)):
        ref = dsp.istft((mag * gains[i][:, None]) * phase)                   # complex128 -> float64 like the reference
        assert (got[i].double() - ref).abs().max().item() < 3e-6, i


def _gpu_mel_db(y2d, n_samples, sumsq=None, ref_rms=0.0):
    cfg = ALPHA_120S
    copies = y2d.shape[0]
    n_frames = 1 + n_samples // 512
    n_cta = -(-n_frames // lib().b200x_mel_frames_per_cta())
    db = torch.full((copies, n_frames, 128), float("nan"), device="cuda")
    cmax = torch.zeros(copies, n_cta, device="cuda")
    ok(lib().b200x_mel_db(P(y2d), y2d.shape[1], n_samples, copies, cfg.sample_rate, cfg.n_mels, cfg.f_min, cfg.f_max, cfg.amin,
                          P(sumsq), ref_rms, n_samples, P(db), n_frames, P(cmax), P(None), 0, P(None)))
    return db, cmax, n_frames, n_cta


def test_mel_db_matches_torchaudio():
    cfg = ALPHA_120S
    ys = np.stack([_track(9.0, "REAL"), 0.05 * _track(9.0, "UDIO")])
    n = ys.shape[1]
    buf = torch.zeros(2, n + 8, device="cuda")
    buf[:, :n] = torch.from_numpy(ys).cuda()
    db, cmax, n_frames, _ = _gpu_mel_db(buf, n)
    power = dsp.mel_frontend(torch.from_numpy(ys), cfg, "power")            # [B, 128, n_frames]
    ref = 10.0 * torch.log10(torch.clamp(power, min=cfg.amin)).transpose(1, 2)
    got = db.cpu()
    # compare where the reference is above its own -80 dB floor (below it only the clamp matters)
    live = ref > ref.amax(dim=(1, 2), keepdim=True) - 80.0
    assert (got - ref)[live].abs().max().item() < 2e-3
    assert torch.allclose(cmax.amax(dim=1).cpu(), ref.amax(dim=(1, 2)), atol=1e-3)


@pytest.mark.parametrize("sample_rate,n_mels,f_min,f_max", [
    (44100, 128, 0.0, 22050.0),      # the reference's shipped sampling rate
    (16000, 64, 20.0, 7600.0),       # fewer, wider filters; band edges inside the spectrum
    (22050, 32, 0.0, 11025.0),       # one filter per lane: the widest segments (every lane walks ~ 33 bins)
    (16000, 96, 0.0, 4000.0),        # f_max at half Nyquist: the upper half of the bins carries no weight
    (8000, 128, 0.0, 4000.0),        # narrow filters: empty segments at the low end (two centre frequencies inside one bin)
])
def test_mel_filterbank_geometries_match_torchaudio(sample_rate, n_mels, f_min, f_max):
    """The filterbank walk (contiguous runs of whole segments per lane, sign-bit flush, compacted non-empty segments) against
    torchaudio's dense HTK filterbank for geometries other than the classifier's own."""
    import dataclasses

    cfg = dataclasses.replace(ALPHA_120S, sample_rate=sample_rate, n_mels=n_mels, f_min=f_min, f_max=f_max)
    ys = np.stack([_track(6.0, "SUNO"), 0.3 * _track(6.0, "REAL")])
    n = ys.shape[1]
    buf = torch.zeros(2, n + 8, device="cuda")
    buf[:, :n] = torch.from_numpy(ys).cuda()
    n_frames = 1 + n // 512
    n_cta = -(-n_frames // lib().b200x_mel_frames_per_cta())
    db = torch.full((2, n_frames, n_mels), float("nan"), device="cuda")
    cmax = torch.zeros(2, n_cta, device="cuda")
    ok(lib().b200x_mel_db(P(buf), buf.shape[1], n, 2, sample_rate, n_mels, f_min, f_max, cfg.amin, P(None), 0.0, n, P(db), n_frames,
                          P(cmax), P(None), 0, P(None)))
    power = dsp.mel_frontend(torch.from_numpy(ys), cfg, "power")
    ref = 10.0 * torch.log10(torch.clamp(power, min=cfg.amin)).transpose(1, 2)
    got = db.cpu()
    assert torch.isfinite(got).all()
    live = ref > ref.amax(dim=(1, 2), keepdim=True) - 80.0
    assert (got - ref)[live].abs().max().item() < 2e-3
    # and back to the classifier's own bank (the per-device table is rebuilt on a geometry change)
    db2, _, _, _ = _gpu_mel_db(buf, n)
    ref2 = 10.0 * torch.log10(torch.clamp(dsp.mel_frontend(torch.from_numpy(ys), ALPHA_120S, "power"), min=ALPHA_120S.amin)).transpose(1, 2)
    live2 = ref2 > ref2.amax(dim=(1, 2), keepdim=True) - 80.0
    assert (db2.cpu() - ref2)[live2].abs().max().item() < 2e-3


def test_mel_normalize_resize_matches_oracle():
    cfg = ALPHA_120S
    ys = np.stack([_track(12.0, "SUNO_PRO"), _track(12.0, "ElevenLabs")])
    n = ys.shape[1]
    buf = torch.zeros(2, n + 8, device="cuda")
    buf[:, :n] = torch.from_numpy(ys).cuda()
    db, cmax, n_frames, n_cta = _gpu_mel_db(buf, n)
    img_t = torch.zeros(2, cfg.input_temp_dim, 128, dtype=torch.bfloat16, device="cuda")
    img_f = torch.zeros(2, 128, cfg.input_temp_dim, dtype=torch.bfloat16, device="cuda")
    partial = torch.zeros(2 * 32 * 2, dtype=torch.float64, device="cuda")
    floor = torch.zeros(2, device="cuda")
    ok(lib().b200x_mel_normalize_resize(P(db), n_frames, P(cmax), n_cta, 2, n_frames, 128, cfg.top_db, 1, cfg.norm_eps,
                                        cfg.input_temp_dim, P(None), P(None), P(None), P(None), P(partial), P(floor), P(img_t),
                                        P(img_f), cfg.input_temp_dim, P(None)))
    ref = spectttra.resize(dsp.mel_frontend(torch.from_numpy(ys), cfg, "norm"), cfg)     # [B, 128, 3744]
    got_f = img_f.float().cpu()
    got_t = img_t.float().cpu().transpose(1, 2)
    assert torch.equal(got_f, got_t)                                          # both operand layouts hold the same image
    assert (got_f - ref).abs().max().item() < 2e-2                            # bf16 storage of O(1) values
    assert (got_f - ref).abs().mean().item() < 2e-3


def test_mel_rms_gain():
    y = _track(5.0, "REAL")
    n = len(y)
    buf = torch.zeros(1, n + 8, device="cuda")
    buf[0, :n] = torch.from_numpy(y).cuda() * 0.5
    ss = torch.tensor([float((0.5 * y.astype(np.float64)) ** 2).sum() if False else float(((0.5 * y.astype(np.float64)) ** 2).sum())],
                      dtype=torch.float64, device="cuda")
    ref_rms = float(np.sqrt(np.mean(y.astype(np.float64) ** 2) + 1e-8))
    db_scaled, _, _, _ = _gpu_mel_db(buf, n, sumsq=ss, ref_rms=ref_rms)
    buf2 = torch.zeros(1, n + 8, device="cuda")
    buf2[0, :n] = torch.from_numpy(dsp.match_rms(y, 0.5 * y).astype(np.float32)).cuda()
    db_ref, _, _, _ = _gpu_mel_db(buf2, n)
    live = db_ref > db_ref.max() - 80
    assert (db_scaled - db_ref)[live].abs().max().item() < 1e-3


def test_mix_stems():
    y, stems = synth.synth_track("UDIO", 1, 16000, 3.0, with_stems=True)
    st = np.stack([stems[k] for k in sorted(stems)])
    masks = np.array([[1, 1, 1, 1], [0, 0, 0, 0], [1, 0, 1, 0], [0, 1, 0, 0]], np.uint8)
    out = torch.full((4, st.shape[1]), float("nan"), device="cuda")
    ok(lib().b200x_mix_stems(P(D(st)), st.shape[1], 4, P(D(masks)), 4, P(out), st.shape[1], P(None)))
    ref = masks.astype(np.float32) @ st
    assert np.abs(out.cpu().numpy() - ref).max() < 1e-6


def test_sparse_occlusion_path_is_bit_identical_to_dense():
    """iSTFT restricted to the samples that classifier frames [t0-4, t1+4) read + baseline spectrogram for every other
    frame == the dense computation, bit for bit (same arithmetic on the same inputs)."""
    cfg = ALPHA_120S
    y = _track(20.0, "UDIO")
    L = len(y)
    S = _gpu_stft(y)
    n_frames = S.shape[0]
    wins = np.array([[0, 40, 0, 200], [3, 50, 10, 61], [4, 70, 0, 1025], [200, 328, 500, 551], [n_frames - 60, n_frames, 900, 1025],
                     [n_frames - 64, n_frames - 4, 0, 51], [100, 100, 0, 10]], np.int32)
    n = len(wins)
    d_w = D(wins)
    # dense reference
    yd = torch.zeros(n, L + 8, device="cuda")
    ok(lib().b200x_istft_masked(P(S), STRIDE, n_frames, n, 1, P(d_w), 0.0, P(None), P(yd), L + 8, P(None), P(None), 0, P(None)))
    db_d, cmax_d, _, n_cta_d = _gpu_mel_db(yd, L)
    img_t_d = torch.zeros(n, cfg.input_temp_dim, 128, dtype=torch.bfloat16, device="cuda")
    img_f_d = torch.zeros(n, 128, cfg.input_temp_dim, dtype=torch.bfloat16, device="cuda")
    part = torch.zeros(n * 64, dtype=torch.float64, device="cuda")
    fl_d = torch.zeros(n, device="cuda")
    ok(lib().b200x_mel_normalize_resize(P(db_d), n_frames, P(cmax_d), n_cta_d, n, n_frames, 128, cfg.top_db, 1, cfg.norm_eps,
                                        cfg.input_temp_dim, P(None), P(None), P(None), P(None), P(part), P(fl_d), P(img_t_d),
                                        P(img_f_d), cfg.input_temp_dim, P(None)))
    # baseline + sparse
    yb = torch.zeros(1, L + 8, device="cuda")
    ok(lib().b200x_istft_masked(P(S), STRIDE, n_frames, 1, 0, P(None), 0.0, P(None), P(yb), L + 8, P(None), P(None), 0, P(None)))
    db_b, _, _, _ = _gpu_mel_db(yb, L)
    pre = torch.zeros(n_frames + 1, device="cuda")
    suf = torch.zeros(n_frames + 1, device="cuda")
    ok(lib().b200x_mel_base_maxima(P(db_b), n_frames, 128, P(pre), P(suf), P(None)))
    assert float(suf[0]) == float(db_b.max()) and float(pre[n_frames]) == float(db_b.max())
    fm = db_b[0, :n_frames].max(dim=1).values.cpu().numpy()                       # per-frame maxima -> running maxima, exactly
    want_pre = np.concatenate([[-np.inf], np.maximum.accumulate(fm)]).astype(np.float32)
    want_suf = np.concatenate([np.maximum.accumulate(fm[::-1])[::-1], [-np.inf]]).astype(np.float32)
    assert np.array_equal(pre.cpu().numpy(), want_pre) and np.array_equal(suf.cpu().numpy(), want_suf)
    rng = torch.zeros(n, 2, dtype=torch.int32, device="cuda")
    ok(lib().b200x_frame_ranges(P(d_w), n, n_frames, P(rng), P(None)))
    max_range = int((wins[:, 1] - wins[:, 0]).max()) + 8
    ys = torch.full((n, L + 8), float("nan"), device="cuda")      # untouched samples must never be read
    ok(lib().b200x_istft_masked(P(S), STRIDE, n_frames, n, 1, P(d_w), 0.0, P(None), P(ys), L + 8, P(None), P(rng), max_range, P(None)))
    n_cta_s = -(-min(n_frames, max_range) // lib().b200x_mel_frames_per_cta())
    db_s = torch.full((n, n_frames, 128), float("nan"), device="cuda")
    cmax_s = torch.zeros(n, n_cta_s, device="cuda")
    ok(lib().b200x_mel_db(P(ys), L + 8, L, n, cfg.sample_rate, 128, cfg.f_min, cfg.f_max, cfg.amin, P(None), 0.0, L, P(db_s), n_frames,
                          P(cmax_s), P(rng), max_range, P(None)))
    img_t_s = torch.zeros_like(img_t_d)
    img_f_s = torch.zeros_like(img_f_d)
    fl_s = torch.zeros(n, device="cuda")
    ok(lib().b200x_mel_normalize_resize(P(db_s), n_frames, P(cmax_s), n_cta_s, n, n_frames, 128, cfg.top_db, 1, cfg.norm_eps,
                                        cfg.input_temp_dim, P(db_b), P(pre), P(suf), P(rng), P(part), P(fl_s), P(img_t_s),
                                        P(img_f_s), cfg.input_temp_dim, P(None)))
    assert torch.equal(fl_s, fl_d)
    assert torch.equal(img_t_s, img_t_d) and torch.equal(img_f_s, img_f_d)
    r = rng.cpu().numpy()
    assert r[-1].tolist() == [0, 0] and r[0].tolist() == [0, 44] and r[4].tolist() == [n_frames - 64, n_frames]
