"""Device polyphase resampler (track-loader front) against scipy.signal.resample_poly with the same filter."""
import numpy as np
import pytest
from scipy.io import wavfile

from audio_deepfake_explainability_b200.audio_io import load_audio, resample_poly_host
from audio_deepfake_explainability_b200.engine import Engine
from audio_deepfake_explainability_b200.weights import ALPHA_120S, random_state_dict

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    e = Engine(ALPHA_120S, random_state_dict(ALPHA_120S, 0), copies_per_chunk=1, max_samples=16000 * 4)
    yield e
    e.close()


@pytest.mark.parametrize("native_sr,sr,n", [(44100, 16000, 44100 * 3 + 17), (22050, 16000, 30001), (8000, 16000, 9000), (48000, 44100, 48000)])
def test_device_resampler_matches_scipy(eng, native_sr, sr, n):
    rng = np.random.default_rng(native_sr + n)
    x = (0.3 * rng.standard_normal(n)).astype(np.float32)
    got = eng.resample(x, native_sr, sr)
    ref = resample_poly_host(x, native_sr, sr)
    assert got.shape == ref.shape and got.dtype == np.float32
    assert np.abs(got - ref).max() <= 1e-6


def test_loader_uses_the_device_resampler(eng, tmp_path):
    t = np.arange(44100 * 2) / 44100.0
    tone = 0.4 * np.sin(2 * np.pi * 440 * t)
    wavfile.write(str(tmp_path / "a.wav"), 44100, (tone * 32767).astype(np.int16))
    y_dev, sr = load_audio(tmp_path / "a.wav", sr=16000, resample=eng.resample)
    y_host, _ = load_audio(tmp_path / "a.wav", sr=16000)
    assert sr == 16000 and len(y_dev) == 32000
    assert np.abs(y_dev - y_host).max() <= 1e-6
    ref = 0.4 * np.sin(2 * np.pi * 440 * np.arange(32000) / 16000.0)
    assert np.abs(y_dev[400:-400] - ref[400:-400]).max() < 2e-4          # the tone survives the rate change
