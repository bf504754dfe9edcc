"""ORACLE (test infrastructure only - never imported by the product path).

CPU restatement of the SpecTTTra-alpha forward pass that the reference reaches through the un-vendored
third-party package ``sonics`` (github awsaf49/sonics, version unpinned; call sites
src/sonics_api.py:20, 246-248, 268-271).  The package is absent from this image and there is no
network, so the published architecture (SONICS paper / public repo: FeatureExtractor -> bilinear resize
-> STTokenizer -> pre-norm ViT blocks -> final LayerNorm -> token mean -> Linear) is restated as a
plain functional forward over a ``sonics``-named state dict.

PARITY STATUS: *unpinned* - no sonics install, checkpoint, golden vector or reference test exists for
this path.  Each unverifiable detail is a field of ``SpecTTTraConfig``.

``gemm_dtype='bf16'`` rounds the inputs of every dense contraction (activations, weights, softmax
probabilities) to bfloat16 and accumulates in float32, i.e. the arithmetic contract of the engine's
tcgen05 kernels; ``'fp32'`` is the reference's own PyTorch arithmetic.
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F

from . import dsp


def _t(sd: Dict[str, np.ndarray], name: str) -> torch.Tensor:
    v = sd[name]
    return v if isinstance(v, torch.Tensor) else torch.from_numpy(np.asarray(v))


def _q(x: torch.Tensor, gemm_dtype: str) -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float32) if gemm_dtype == "bf16" else x


def _linear(x, w, b, gemm_dtype):
    y = _q(x, gemm_dtype) @ _q(w, gemm_dtype).transpose(-1, -2)
    return y if b is None else y + b


def resize(spec: torch.Tensor, cfg) -> torch.Tensor:
    """``F.interpolate(spec[:, None], size=input_shape, mode='bilinear')`` -> ``[B, F, T]``."""
    out = F.interpolate(spec.unsqueeze(1), size=(cfg.input_spec_dim, cfg.input_temp_dim), mode="bilinear")
    return out.squeeze(1)


def tokenize(img: torch.Tensor, sd, cfg, gemm_dtype: str = "fp32") -> torch.Tensor:
    """sonics STTokenizer: temporal Conv1d(F->D, k=t_clip, s=t_clip) over time and spectral
    Conv1d(T->D, k=f_clip, s=f_clip) over frequency, each + GELU + positional encoding (+ LayerNorm when
    ``pre_norm``), concatenated ``[temporal, spectral]`` -> ``[B, n_tokens, D]``."""
    outs = []
    for name, x, clip in (
        ("temporal_tokenizer", img, cfg.t_clip),
        ("spectral_tokenizer", img.transpose(1, 2), cfg.f_clip),
    ):
        p = "encoder.st_tokenizer." + name + "."
        w = _t(sd, p + "conv1d.weight")                     # [D, Cin, clip]
        b = _t(sd, p + "conv1d.bias") if (p + "conv1d.bias") in sd else None
        n_tok = (x.shape[2] - clip) // clip + 1
        # conv1d with stride == kernel == clip is a matmul over unfolded clips
        patches = x[:, :, : n_tok * clip].reshape(x.shape[0], x.shape[1], n_tok, clip)   # [B, Cin, n, clip]
        patches = patches.permute(0, 2, 1, 3).reshape(x.shape[0], n_tok, -1)             # [B, n, Cin*clip]
        t = _linear(patches, w.reshape(w.shape[0], -1), b, gemm_dtype)
        t = F.gelu(t)
        if cfg.pe_learnable:
            t = t + _t(sd, p + "pos_encoder.pe")[:n_tok]
        else:
            t = t + sinusoid_pe(n_tok, cfg.embed_dim)
        if cfg.pre_norm:
            t = F.layer_norm(t, (cfg.embed_dim,), _t(sd, p + "norm_pre.weight"), _t(sd, p + "norm_pre.bias"),
                             cfg.tokenizer_ln_eps)
        outs.append(t)
    return torch.cat(outs, dim=1)


def sinusoid_pe(n: int, d: int) -> torch.Tensor:
    pos = torch.arange(n, dtype=torch.float32).unsqueeze(1)
    div = torch.exp(torch.arange(0, d, 2, dtype=torch.float32) * (-np.log(10000.0) / d))
    pe = torch.zeros(n, d)
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe


def attention(x, sd, p, cfg, gemm_dtype):
    B, N, D = x.shape
    H, hd = cfg.num_heads, cfg.head_dim
    qkv = _linear(x, _t(sd, p + "attn.qkv.weight"), _t(sd, p + "attn.qkv.bias") if cfg.qkv_bias else None, gemm_dtype)
    qkv = _q(qkv, gemm_dtype)                                # engine stores qkv in the GEMM input dtype
    qkv = qkv.reshape(B, N, 3, H, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    s = (q @ k.transpose(-1, -2)) * (hd ** -0.5)
    pattn = torch.softmax(s, dim=-1)
    if gemm_dtype == "bf16":
        # engine: P = exp(s - m) in bf16 for the PV contraction, normalised by the row sum (of the same bf16 values)
        m = s.max(dim=-1, keepdim=True).values
        e = torch.exp(s - m)
        eq = _q(e, gemm_dtype)
        o = (eq @ v) / eq.sum(dim=-1, keepdim=True)
    else:
        o = pattn @ v
    o = o.transpose(1, 2).reshape(B, N, D)
    return _linear(o, _t(sd, p + "attn.proj.weight"), _t(sd, p + "attn.proj.bias"), gemm_dtype)


def encoder(tokens: torch.Tensor, sd, cfg, gemm_dtype: str = "fp32", return_all: bool = False):
    x = tokens
    trace = []
    for i in range(cfg.num_layers):
        p = f"encoder.transformer.blocks.{i}."
        h = F.layer_norm(x, (cfg.embed_dim,), _t(sd, p + "norm1.weight"), _t(sd, p + "norm1.bias"), cfg.block_ln_eps)
        x = x + attention(h, sd, p, cfg, gemm_dtype)
        h = F.layer_norm(x, (cfg.embed_dim,), _t(sd, p + "norm2.weight"), _t(sd, p + "norm2.bias"), cfg.block_ln_eps)
        h = F.gelu(_linear(h, _t(sd, p + "mlp.fc1.weight"), _t(sd, p + "mlp.fc1.bias"), gemm_dtype))
        x = x + _linear(h, _t(sd, p + "mlp.fc2.weight"), _t(sd, p + "mlp.fc2.bias"), gemm_dtype)
        if return_all:
            trace.append(x)
    if cfg.final_norm:
        x = F.layer_norm(x, (cfg.embed_dim,), _t(sd, "encoder.transformer.norm.weight"),
                         _t(sd, "encoder.transformer.norm.bias"), cfg.block_ln_eps)
    return (x, trace) if return_all else x


def head(features: torch.Tensor, sd) -> torch.Tensor:
    """``features.mean(dim=1)`` -> ``Linear(D, 1)`` -> logit ``[B]`` (fp32 in the engine too)."""
    emb = features.mean(dim=1)
    return (emb @ _t(sd, "classifier.weight").transpose(0, 1) + _t(sd, "classifier.bias")).squeeze(-1)


@torch.no_grad()
def forward_logits(audio, sd, cfg, gemm_dtype: str = "fp32") -> torch.Tensor:
    """HFAudioClassifier.forward restated: audio ``[B, L]`` -> logits ``[B]``."""
    audio = torch.as_tensor(np.asarray(audio), dtype=torch.float32) if not isinstance(audio, torch.Tensor) else audio
    if audio.dim() == 1:
        audio = audio.unsqueeze(0)
    spec = dsp.mel_frontend(audio, cfg)
    img = resize(spec, cfg)
    tok = tokenize(img, sd, cfg, gemm_dtype)
    feats = encoder(tok, sd, cfg, gemm_dtype)
    return head(feats, sd)


class OraclePredictor:
    """Duck-typed predictor (``predict(wave, sr) -> float``) = LocalSonnics.predict restated
    (src/sonics_api.py:259-271): ``sigmoid(model(tensor(wave).float()[None])).item()``; ``sr`` ignored."""

    def __init__(self, sd, cfg, gemm_dtype: str = "fp32"):
        self.sd, self.cfg, self.gemm_dtype = sd, cfg, gemm_dtype
        self.calls = 0

    def predict(self, audio_wave, sr: int) -> float:
        self.calls += 1
        t = torch.tensor(np.asarray(audio_wave)).float().unsqueeze(0)
        return torch.sigmoid(forward_logits(t, self.sd, self.cfg, self.gemm_dtype)).item()
