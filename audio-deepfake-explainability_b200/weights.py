"""SpecTTTra-alpha classifier configuration and seeded random-init weights.

The classifier arithmetic is not in /root/reference: it is the un-vendored third-party package
``sonics`` (``from sonics import HFAudioClassifier``, src/sonics_api.py:20,246-248), HF weights
``awsaf49/sonics-spectttra-alpha-120s`` (configs/Spec_occlusion_configs/spectrogram_explainability.yaml:19).
No checkpoint can be fetched here (no network), so both the engine and the test oracle use the same
seeded random-init state dict, keyed with ``sonics``-compatible parameter names so a real checkpoint
loads through the same path.  Every structural detail that could not be verified offline is a named
field of :class:`SpecTTTraConfig`.
"""
from __future__ import annotations

from dataclasses import dataclass, asdict
from typing import Dict

import numpy as np


@dataclass(frozen=True)
class SpecTTTraConfig:
    # audio / mel front-end (torchaudio MelSpectrogram + AmplitudeToDB as used by sonics)
    sample_rate: int = 16000
    n_fft: int = 2048
    hop_length: int = 512
    win_length: int = 2048
    n_mels: int = 128
    f_min: float = 20.0
    f_max: float = 8000.0
    top_db: float = 80.0
    amin: float = 1e-10
    norm_eps: float = 1e-6          # (x - mean) / (std + eps)
    std_unbiased: bool = True       # torch.std default
    # encoder
    input_spec_dim: int = 128
    input_temp_dim: int = 3744
    t_clip: int = 3
    f_clip: int = 1
    embed_dim: int = 384
    num_heads: int = 6
    num_layers: int = 12
    mlp_ratio: float = 2.67
    pre_norm: bool = True           # tokenizer conv has no bias, LayerNorm(eps=1e-6) after PE
    pe_learnable: bool = True
    qkv_bias: bool = False
    tokenizer_ln_eps: float = 1e-6
    block_ln_eps: float = 1e-5
    final_norm: bool = True
    num_classes: int = 1

    @property
    def head_dim(self) -> int:
        return self.embed_dim // self.num_heads

    @property
    def mlp_hidden(self) -> int:
        return int(self.embed_dim * self.mlp_ratio)

    @property
    def num_temporal_tokens(self) -> int:
        return (self.input_temp_dim - self.t_clip) // self.t_clip + 1

    @property
    def num_spectral_tokens(self) -> int:
        return (self.input_spec_dim - self.f_clip) // self.f_clip + 1

    @property
    def num_tokens(self) -> int:
        return self.num_temporal_tokens + self.num_spectral_tokens

    def to_dict(self) -> dict:
        return asdict(self)


ALPHA_120S = SpecTTTraConfig()


def tiny_config(**kw) -> SpecTTTraConfig:
    """A small configuration of the same architecture for fast CPU tests."""
    base = dict(input_temp_dim=96, embed_dim=64, num_heads=2, num_layers=2, mlp_ratio=2.67,
                n_fft=256, hop_length=64, win_length=256, n_mels=32, input_spec_dim=32)
    base.update(kw)
    return SpecTTTraConfig(**base)


def random_state_dict(cfg: SpecTTTraConfig = ALPHA_120S, seed: int = 0, head_gain: float = 1.0) -> Dict[str, np.ndarray]:
    """Seeded random-init parameters (float32 numpy) with sonics-style names.

    Init mirrors the default ``torch.nn`` initialisers (uniform +-1/sqrt(fan_in) for Linear/Conv
    weights and biases, ones/zeros for LayerNorm, ``randn*0.02`` for the learned positional encodings),
    drawn from ``numpy.random.default_rng(seed)`` so it is reproducible without torch.
    ``head_gain`` scales the classifier weight so that random-init delta-probabilities are not vanishingly
    small (hazards).
    """
    rng = np.random.default_rng(seed)
    D, H = cfg.embed_dim, cfg.mlp_hidden
    sd: Dict[str, np.ndarray] = {}

    def uni(shape, fan_in):
        b = 1.0 / np.sqrt(fan_in)
        return rng.uniform(-b, b, size=shape).astype(np.float32)

    tok = "encoder.st_tokenizer."
    for name, cin, clip, ntok in (
        ("temporal_tokenizer", cfg.input_spec_dim, cfg.t_clip, cfg.num_temporal_tokens),
        ("spectral_tokenizer", cfg.input_temp_dim, cfg.f_clip, cfg.num_spectral_tokens),
    ):
        p = tok + name + "."
        sd[p + "conv1d.weight"] = uni((D, cin, clip), cin * clip)
        if not cfg.pre_norm:
            sd[p + "conv1d.bias"] = uni((D,), cin * clip)
        if cfg.pe_learnable:
            sd[p + "pos_encoder.pe"] = (rng.standard_normal((ntok, D)) * 0.02).astype(np.float32)
        if cfg.pre_norm:
            sd[p + "norm_pre.weight"] = np.ones(D, np.float32)
            sd[p + "norm_pre.bias"] = np.zeros(D, np.float32)
    for i in range(cfg.num_layers):
        p = f"encoder.transformer.blocks.{i}."
        # LayerNorm affine parameters are perturbed from (1, 0) so that parity tests exercise them.
        sd[p + "norm1.weight"] = (1.0 + 0.1 * rng.standard_normal(D)).astype(np.float32)
        sd[p + "norm1.bias"] = (0.05 * rng.standard_normal(D)).astype(np.float32)
        sd[p + "attn.qkv.weight"] = uni((3 * D, D), D)
        if cfg.qkv_bias:
            sd[p + "attn.qkv.bias"] = uni((3 * D,), D)
        sd[p + "attn.proj.weight"] = uni((D, D), D)
        sd[p + "attn.proj.bias"] = uni((D,), D)
        sd[p + "norm2.weight"] = (1.0 + 0.1 * rng.standard_normal(D)).astype(np.float32)
        sd[p + "norm2.bias"] = (0.05 * rng.standard_normal(D)).astype(np.float32)
        sd[p + "mlp.fc1.weight"] = uni((H, D), D)
        sd[p + "mlp.fc1.bias"] = uni((H,), D)
        sd[p + "mlp.fc2.weight"] = uni((D, H), H)
        sd[p + "mlp.fc2.bias"] = uni((D,), H)
    if cfg.final_norm:
        sd["encoder.transformer.norm.weight"] = np.ones(D, np.float32)
        sd["encoder.transformer.norm.bias"] = np.zeros(D, np.float32)
    sd["classifier.weight"] = (uni((cfg.num_classes, D), D) * head_gain).astype(np.float32)
    sd["classifier.bias"] = uni((cfg.num_classes,), D)
    return sd
