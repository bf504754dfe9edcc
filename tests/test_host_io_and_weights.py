"""Host-side pieces either side of the hot path: the track loader that stands in for ``librosa.load(path, sr, duration,
mono=True)`` (src/spectrogram_explainability.py:601, src/dsp_band_ops.py:679), the WAV writer, and the SpecTTTra-alpha-120s
configuration / state-dict layout."""
import numpy as np
import pytest
from scipy.io import wavfile

from audio_deepfake_explainability_b200 import grid
from audio_deepfake_explainability_b200.audio_io import load_audio, write_wav
from audio_deepfake_explainability_b200.weights import ALPHA_120S, random_state_dict
from oracle import loops


def test_pcm16_wav_round_trip_and_duration(tmp_path):
    y = (0.3 * np.sin(2 * np.pi * 440 * np.arange(32000) / 16000)).astype(np.float32)
    write_wav(tmp_path / "a.wav", y, 16000)
    rate, raw = wavfile.read(str(tmp_path / "a.wav"))
    assert rate == 16000 and raw.dtype == np.int16                     # soundfile.write(path, y, sr) default subtype: PCM_16 (:494)
    want = np.clip(np.rint(y.astype(np.float64) * 32768.0), -32768, 32767) / 32768.0
    back, sr = load_audio(tmp_path / "a.wav", sr=16000)
    assert sr == 16000 and back.dtype == np.float32 and np.array_equal(back, want.astype(np.float32))
    assert np.abs(back - y).max() <= 0.5 / 32768 + 1e-9
    write_wav(tmp_path / "clip.wav", np.array([1.5, -1.5, 1.0, -1.0]), 16000)
    assert wavfile.read(str(tmp_path / "clip.wav"))[1].tolist() == [32767, -32768, 32767, -32768]
    cut, _ = load_audio(tmp_path / "a.wav", sr=16000, duration=0.5)
    assert np.array_equal(cut, back[:8000])                            # duration trims BEFORE resampling, like librosa
    native, sr = load_audio(tmp_path / "a.wav", sr=None)
    assert sr == 16000 and len(native) == 32000


def test_int16_stereo_downmix_and_resample(tmp_path):
    t = np.arange(44100) / 44100.0
    left = 0.5 * np.sin(2 * np.pi * 220 * t)
    right = 0.25 * np.sin(2 * np.pi * 220 * t)
    pcm = np.stack([left, right], axis=1)
    wavfile.write(str(tmp_path / "s.wav"), 44100, (pcm * 32767).astype(np.int16))
    y, sr = load_audio(tmp_path / "s.wav", sr=None)
    assert sr == 44100 and y.shape == (44100,)
    assert np.abs(y - 0.375 * np.sin(2 * np.pi * 220 * t)).max() < 2e-4  # channel mean, int16 scaled by 1/32768
    y16, sr = load_audio(tmp_path / "s.wav", sr=16000)
    assert sr == 16000 and len(y16) == 16000
    ref = 0.375 * np.sin(2 * np.pi * 220 * np.arange(16000) / 16000.0)
    assert np.abs(y16[200:-200] - ref[200:-200]).max() < 2e-3            # polyphase resampler: same tone at the new rate


def test_unsupported_container_is_an_error(tmp_path):
    from audio_deepfake_explainability_b200.audio_io import AudioDecodeError
    (tmp_path / "x.mp3").write_bytes(b"ID3")
    with pytest.raises(AudioDecodeError):
        load_audio(tmp_path / "x.mp3")
    (tmp_path / "bad.wav").write_bytes(b"not a wav file")
    with pytest.raises(AudioDecodeError):
        load_audio(tmp_path / "bad.wav")


def test_alpha_120s_configuration_and_state_dict_layout():
    c = ALPHA_120S
    assert (c.num_temporal_tokens, c.num_spectral_tokens, c.num_tokens) == (1248, 128, 1376)
    assert (c.embed_dim, c.num_heads, c.head_dim, c.num_layers, c.mlp_hidden) == (384, 6, 64, 12, 1025)
    sd = random_state_dict(c, 0)
    n_params = sum(int(np.prod(v.shape)) for v in sd.values())
    assert 18.0e6 < n_params < 19.5e6                                    # paper: ~19 M parameters
    assert sd["encoder.transformer.blocks.0.attn.qkv.weight"].shape == (1152, 384)
    assert sd["encoder.transformer.blocks.11.mlp.fc1.weight"].shape == (1025, 384)
    assert sd["encoder.transformer.blocks.11.mlp.fc2.weight"].shape == (384, 1025)
    assert sd["classifier.weight"].shape == (1, 384)
    assert all(k.startswith(("encoder.", "classifier", "ft_extractor")) for k in sd)
    again = random_state_dict(c, 0)
    assert all(np.array_equal(sd[k], again[k]) for k in sd)              # seeded: engine and oracle load the same numbers
    assert not np.array_equal(sd["classifier.weight"], random_state_dict(c, 1)["classifier.weight"])


def test_topk_groups_match_the_oracle_on_ties():
    imps = [0.2, -0.2, 0.0, 0.2, -0.5, 0.05, -0.05, 0.5]
    patches = [{"t_start": i, "t_end": i + 1, "f_start": 0, "f_end": 1, "importance": v} for i, v in enumerate(imps)]
    want = loops.top_window_groups(patches, 3, "f", 512, 16000)
    got = grid.topk_window_groups(imps, 3)
    for g in ("all", "best", "worst", "most_influential"):
        assert [w["t_start"] for w in want[g]["windows"]] == [int(i) for i in got[g]], g
