"""Fused tcgen05 attention vs fp32 softmax attention on the same bf16 qkv (through the C ABI)."""
import pytest
import torch

from gpu_util import D, P, lib, ok

pytestmark = pytest.mark.gpu


def _ref(qkv, copies, tokens, heads):
    x = qkv.float().reshape(copies, tokens, 3, heads, 64).permute(2, 0, 3, 1, 4)
    q, k, v = x[0], x[1], x[2]
    s = (q @ k.transpose(-1, -2)) * 0.125
    o = torch.softmax(s, -1) @ v
    return o.transpose(1, 2).reshape(copies * tokens, heads * 64)


@pytest.mark.parametrize("copies,tokens,heads,scale", [
    (1, 128, 1, 1.0), (1, 64, 2, 1.0), (2, 272, 2, 1.0), (2, 1376, 6, 1.0), (1, 1376, 6, 6.0),
])
@pytest.mark.parametrize("tiles_per_cta,variant", [(5, 256), (6, 256), (4, 256), (4, 0), (2, 65792), (3, 256), (3, 0), (0, 256), (0, 0), (1, 256), (1, 0), (2, 256), (2, 0)])
def test_attention_matches_reference(copies, tokens, heads, scale, tiles_per_cta, variant):
    # tiles_per_cta 3 = production kernel (one tile per CTA, software-pipelined softmax loop); 0 = split-row kernel (one tile per CTA, two threads per row); variant 256 = a quarter of the exponentials
    # as an FMA-pipe polynomial
    import ctypes
    lib().b200x_debug_attention_tiles_per_cta(ctypes.c_int(tiles_per_cta))
    lib().b200x_debug_attention_variant(ctypes.c_int(variant))
    g = torch.Generator(device="cpu").manual_seed(tokens + heads)
    qkv = (torch.randn(copies * tokens, 3 * heads * 64, generator=g) * scale).to(torch.bfloat16)
    ref = _ref(qkv, copies, tokens, heads)
    out = torch.full((copies * tokens, heads * 64), float("nan"), dtype=torch.bfloat16, device="cuda")
    try:
        ok(lib().b200x_attention(P(D(qkv)), P(out), copies, tokens, heads, 64, P(None)))
        torch.cuda.synchronize()
    finally:
        lib().b200x_debug_attention_tiles_per_cta(ctypes.c_int(1))
        lib().b200x_debug_attention_variant(ctypes.c_int(256))
    got = out.float().cpu()
    assert torch.isfinite(got).all()
    err = (got - ref).abs().max().item()
    assert err < 2e-2 * max(1.0, ref.abs().max().item()), f"max err {err}"


def test_attention_rejects_bad_shapes():
    q = torch.zeros(100, 192, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(RuntimeError):
        ok(lib().b200x_attention(P(q), P(q), 1, 100, 1, 64, P(None)))      # tokens % 16 != 0
    with pytest.raises(RuntimeError):
        ok(lib().b200x_attention(P(q), P(q), 1, 96, 1, 32, P(None)))       # head_dim != 64
