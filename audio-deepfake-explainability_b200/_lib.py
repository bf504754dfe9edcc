"""ctypes binding of ``libb200xai.so`` (the C ABI declared in ``include/b200xai.h``).

Fails loudly: if the shared library is missing or a call returns a non-zero status a ``RuntimeError`` is
raised - there is no CPU fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

HERE = Path(__file__).resolve().parent
# B200X_LIB_PATH: developer override for A/B runs of kernel-variant builds (build.py B200X_LIB_OUT); default = the in-tree library
LIB_PATH = Path(os.environ["B200X_LIB_PATH"]) if os.environ.get("B200X_LIB_PATH") else HERE / "libb200xai.so"

c_i32p = C.POINTER(C.c_int32)
c_f32p = C.POINTER(C.c_float)
c_f64p = C.POINTER(C.c_double)
c_u8p = C.POINTER(C.c_uint8)
VP = C.c_void_p


class ModelConfig(C.Structure):
    _fields_ = [
        ("sample_rate", C.c_int32), ("n_fft", C.c_int32), ("hop_length", C.c_int32), ("n_mels", C.c_int32),
        ("f_min", C.c_double), ("f_max", C.c_double), ("top_db", C.c_double), ("amin", C.c_double),
        ("norm_eps", C.c_float), ("std_unbiased", C.c_int32),
        ("input_spec_dim", C.c_int32), ("input_temp_dim", C.c_int32), ("t_clip", C.c_int32), ("f_clip", C.c_int32),
        ("embed_dim", C.c_int32), ("num_heads", C.c_int32), ("num_layers", C.c_int32), ("mlp_hidden", C.c_int32),
        ("pre_norm", C.c_int32), ("pe_learnable", C.c_int32), ("qkv_bias", C.c_int32), ("final_norm", C.c_int32),
        ("tokenizer_ln_eps", C.c_float), ("block_ln_eps", C.c_float),
    ]


# name -> (restype, argtypes); must list every symbol declared in include/b200xai.h
SIGNATURES = {
    "b200x_last_error": (C.c_char_p, []),
    "b200x_version": (C.c_int, []),
    "b200x_set_device": (C.c_int, [C.c_int]),
    "b200x_device_count": (C.c_int, [c_i32p]),
    "b200x_stft": (C.c_int, [VP, C.c_int64, C.c_int, C.c_int, C.c_int, VP, C.c_int, VP]),
    "b200x_istft_masked": (C.c_int, [VP, C.c_int, C.c_int, C.c_int, C.c_int, VP, C.c_float, VP, VP, C.c_int64, VP, VP, C.c_int, VP]),
    "b200x_frame_ranges": (C.c_int, [VP, C.c_int, C.c_int, VP, VP]),
    "b200x_mel_base_maxima": (C.c_int, [VP, C.c_int, C.c_int, VP, VP, VP]),
    "b200x_mel_frames_per_cta": (C.c_int, []),
    "b200x_mel_db": (C.c_int, [VP, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                               VP, C.c_double, C.c_int64, VP, C.c_int, VP, VP, C.c_int, VP]),
    "b200x_mel_normalize_resize": (C.c_int, [VP, C.c_int, VP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int,
                                             C.c_float, C.c_int, VP, VP, VP, VP, VP, VP, VP, VP, C.c_int, VP]),
    "b200x_mix_stems": (C.c_int, [VP, C.c_int64, C.c_int, VP, C.c_int, VP, C.c_int64, VP]),
    "b200x_gemm_bf16": (C.c_int, [VP, C.c_int, VP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, VP, C.c_int, C.c_int,
                                  VP, C.c_int, VP, VP, C.c_int, C.c_int, C.c_int, C.c_int, VP]),
    "b200x_gemm_tokens_mmajor": (C.c_int, [VP, C.c_int, C.c_int, VP, C.c_int, C.c_int, VP, C.c_int, VP, C.c_int, VP, C.c_int, C.c_int, VP]),
    "b200x_gemm_bf16_astationary": (C.c_int, [VP, C.c_int, VP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, VP, C.c_int, VP, C.c_int, C.c_int, VP]),
    "b200x_gemm_resid_ln_bf16": (C.c_int, [VP, C.c_int, VP, C.c_int, C.c_int, C.c_int, C.c_int, VP, C.c_int, VP, VP, VP, C.c_float, VP, C.c_int,
                                           C.c_int, VP]),
    "b200x_attention": (C.c_int, [VP, VP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, VP]),
    "b200x_layernorm": (C.c_int, [VP, C.c_int, C.c_int, VP, VP, VP, VP, C.c_int, C.c_int, C.c_float, VP, VP, C.c_int, VP]),
    "b200x_head_slices": (C.c_int, []),
    "b200x_head": (C.c_int, [VP, C.c_int, C.c_int, C.c_int, VP, VP, C.c_float, C.c_int, VP, C.c_float, VP, VP, VP, VP]),
    "b200x_delta": (C.c_int, [VP, C.c_float, C.c_int, VP, VP]),
    "b200x_delta_dev": (C.c_int, [VP, VP, C.c_int, VP, VP]),
    "b200x_saliency_reduce": (C.c_int, [VP, VP, C.c_int, C.c_int, C.c_int, VP, VP]),
    "b200x_istft_masked_tracks": (C.c_int, [VP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int, VP, C.c_float, VP, VP, C.c_int64, VP, VP,
                                             C.c_int, VP]),
    "b200x_mel_db_ref": (C.c_int, [VP, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, VP, C.c_double, VP,
                                    C.c_int64, VP, C.c_int, VP, VP, C.c_int, VP]),
    "b200x_wave_rms": (C.c_int, [VP, C.c_int64, C.c_int64, C.c_int, C.c_int, VP, VP]),
    "b200x_stft_batch": (C.c_int, [VP, C.c_int64, C.c_int64, C.c_int, VP, C.c_int, C.c_int64, VP]),
    "b200x_mel_power": (C.c_int, [VP, C.c_int, C.c_int, C.c_int, VP, VP, VP, VP]),
    "b200x_mel_nnls": (C.c_int, [VP, C.c_int, C.c_int, C.c_int, C.c_int, VP, C.c_float, VP, VP, VP, VP, VP, VP, C.c_float, C.c_int, VP, C.c_int,
                                 VP, C.c_int64, VP]),
    "b200x_gl_init": (C.c_int, [VP, C.c_int64, VP, C.c_int64, C.c_int, C.c_int, C.c_uint32, C.c_int, VP]),
    "b200x_gl_update": (C.c_int, [VP, VP, VP, C.c_int64, VP, C.c_int64, C.c_int, C.c_int, C.c_float, VP]),
    "b200x_engine_saliency_map_shape": (C.c_int, [VP, VP, VP, C.c_int, C.c_int, C.c_int, VP]),
    "b200x_engine_set_mel_basis": (C.c_int, [VP, C.c_int, VP, VP, C.c_float]),
    "b200x_engine_mel_spectrogram": (C.c_int, [VP, VP]),
    "b200x_engine_mel_sweep": (C.c_int, [VP, C.c_int, VP, VP, C.c_int, C.c_float, C.c_int, C.c_int, C.c_uint32, C.c_int, C.c_float, VP, VP]),
    "b200x_resample_poly": (C.c_int, [VP, C.c_int64, C.c_int, C.c_int, VP, C.c_int, VP, C.c_int64, VP]),
    "b200x_engine_resample": (C.c_int, [VP, VP, C.c_int64, C.c_int, C.c_int, VP, C.c_int, VP, C.c_int64]),
    "b200x_rise_map": (C.c_int, [VP, C.c_int, C.c_uint32, C.c_double, C.c_int, C.c_int, VP, VP]),
    "b200x_band_map": (C.c_int, [VP, VP, C.c_int, C.c_int, C.c_int, VP, VP]),
    "b200x_rank": (C.c_int, [VP, C.c_int, C.c_int, VP, VP]),
    "b200x_engine_create": (C.c_int, [C.POINTER(ModelConfig), C.c_int, C.c_int64, C.POINTER(VP)]),
    "b200x_engine_destroy": (None, [VP]),
    "b200x_engine_set_param": (C.c_int, [VP, C.c_char_p, VP, C.c_int64]),
    "b200x_engine_finalize": (C.c_int, [VP]),
    "b200x_engine_predict": (C.c_int, [VP, VP, C.c_int64, C.c_int, C.c_int, VP, VP]),
    "b200x_engine_set_track": (C.c_int, [VP, VP, C.c_int64, C.c_int]),
    "b200x_engine_track_shape": (C.c_int, [VP, c_i32p, c_i32p]),
    "b200x_engine_get_spectrogram": (C.c_int, [VP, VP]),
    "b200x_engine_occlusion_sweep": (C.c_int, [VP, VP, C.c_int, C.c_float, C.c_int, VP]),
    "b200x_engine_fbp_sweep": (C.c_int, [VP, VP, C.c_int, C.c_int, C.c_int, VP]),
    "b200x_engine_stem_sweep": (C.c_int, [VP, VP, C.c_int, C.c_int64, VP, C.c_int, C.c_int, VP]),
    "b200x_engine_window_audio": (C.c_int, [VP, VP, C.c_int, VP, C.c_int64, VP]),
    "b200x_engine_occluded_audio": (C.c_int, [VP, VP, C.c_int, C.c_float, VP]),
    "b200x_engine_band_audio": (C.c_int, [VP, VP, C.c_int, VP]),
    "b200x_engine_predict_track": (C.c_int, [VP, VP, VP]),
    "b200x_engine_occlusion_sweep_base": (C.c_int, [VP, VP, C.c_int, C.c_float, C.c_int, VP, VP]),
    "b200x_engine_set_alternate": (C.c_int, [VP, C.c_int]),
    "b200x_engine_fbp_sweep_tracks": (C.c_int, [VP, VP, C.c_int, C.c_int64, VP, C.c_int, C.c_int, VP, VP]),
    "b200x_engine_saliency_map": (C.c_int, [VP, VP, VP, C.c_int, VP]),
    "b200x_engine_rise_sweep": (C.c_int, [VP, C.c_int, C.c_int, C.c_uint32, C.c_double, C.c_int, VP]),
    "b200x_engine_rise_audio": (C.c_int, [VP, C.c_int, C.c_int, C.c_uint32, C.c_double, VP]),
    "b200x_engine_rise_map": (C.c_int, [VP, VP, C.c_int, C.c_uint32, C.c_double, VP]),
    "b200x_engine_band_map": (C.c_int, [VP, VP, VP, C.c_int, VP]),
    "b200x_engine_rank": (C.c_int, [VP, VP, C.c_int, C.c_int, VP]),
    "b200x_engine_debug_buffer": (C.c_int, [VP, C.c_char_p, C.POINTER(VP), C.POINTER(C.c_int64)]),
    "b200x_engine_set_trace": (C.c_int, [VP, VP]),
    "b200x_engine_launch_count": (C.c_int64, [VP]),
    "b200x_engine_set_timing": (C.c_int, [VP, C.c_int]),
    "b200x_engine_set_graphs": (C.c_int, [VP, C.c_int]),
    "b200x_engine_set_fused_layernorm": (C.c_int, [VP, C.c_int]),
    "b200x_engine_get_timing": (C.c_int, [VP, VP, VP]),
    "b200x_engine_stream": (VP, [VP]),
    "b200x_engine_synchronize": (C.c_int, [VP]),
}

_lib = None


def load():
    """Load the shared library and bind every symbol; raises RuntimeError if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` (nvcc, sm_100a). "
            "This engine has no CPU fallback."
        )
    lib = C.CDLL(str(LIB_PATH), mode=getattr(os, "RTLD_NOW", 2))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the library does not export the symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = load().b200x_last_error()
        raise RuntimeError(f"b200xai {what} failed (status {status}): {msg.decode('utf-8', 'replace') if msg else ''}")
