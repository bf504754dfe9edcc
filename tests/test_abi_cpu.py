"""CPU-side checks of the C ABI: the library loads without a GPU and exports every symbol that
include/b200xai.h declares; the ctypes table covers the header; compute entry points are not called here."""
import ctypes
import re
from pathlib import Path

from audio_deepfake_explainability_b200 import _lib

ROOT = Path(__file__).resolve().parents[1]


def _header_symbols():
    text = (ROOT / "include" / "b200xai.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200x_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    syms = _header_symbols()
    assert len(syms) >= 40
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/b200xai.h but not exported"


def test_ctypes_table_matches_header():
    assert sorted(_lib.SIGNATURES) == _header_symbols()


def test_load_binds_and_reports_version():
    lib = _lib.load()
    assert lib.b200x_version() == 100
    assert lib.b200x_last_error() is not None
    assert lib.b200x_mel_frames_per_cta() > 0 and lib.b200x_head_slices() > 0


def test_model_config_struct_layout_matches_header():
    # 4 int32, 4 double, float, int32, 4 int32, 4 int32, 4 int32, 2 float  (natural alignment, no packing)
    assert ctypes.sizeof(_lib.ModelConfig) == 16 + 32 + 8 + 16 + 16 + 16 + 8
