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
    (1, 128, 1, 1.0), (1, 64, 2, 1.0), (2, 272, 2, 1.0), (2, 1376, 6, 1.0), (1, 1376, 6, 6.0), (3, 1376, 6, 1.0),
])
@pytest.mark.parametrize("reverse", [0, 1])
def test_attention_matches_reference(copies, tokens, heads, scale, reverse):
    # scale 6.0 drives the lazy-rescaling path (row maxima that grow by more than 2^8 between key tiles)
    g = torch.Generator(device="cpu").manual_seed(tokens + heads)
    qkv = (torch.randn(copies * tokens, 3 * heads * 64, generator=g) * scale).to(torch.bfloat16)
    ref = _ref(qkv, copies, tokens, heads)
    out = torch.full((copies * tokens, heads * 64), float("nan"), dtype=torch.bfloat16, device="cuda")
    ok(lib().b200x_attention(P(D(qkv)), P(out), copies, tokens, heads, 64, reverse, P(None)))
    torch.cuda.synchronize()
    got = out.float().cpu()
    assert torch.isfinite(got).all()
    err = (got - ref).abs().max().item()
    assert err < 2e-2 * max(1.0, ref.abs().max().item()), f"max err {err}"


def test_attention_rejects_bad_shapes():
    q = torch.zeros(100, 192, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(RuntimeError):
        ok(lib().b200x_attention(P(q), P(q), 1, 100, 1, 64, 0, P(None)))      # tokens % 16 != 0
    with pytest.raises(RuntimeError):
        ok(lib().b200x_attention(P(q), P(q), 1, 96, 1, 32, 0, P(None)))       # head_dim != 64
