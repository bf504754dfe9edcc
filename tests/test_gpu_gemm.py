"""tcgen05 GEMM kernel vs a plain fp32 matmul of the same bf16-rounded operands (through the C ABI)."""
import pytest
import torch

from gpu_util import D, P, bf16_round, lib, ok

pytestmark = pytest.mark.gpu
OUT_BF16, OUT_RESID, OUT_TOKEN = 0, 1, 2


def _run(M, N, K, bn, mode, bias=True, gelu=False, group=None, lda=None, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    lda = lda or K
    a = torch.randn(M, lda, generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(torch.bfloat16)
    b = torch.randn(N, generator=g) if bias else None
    ref = a[:, :K].float() @ w.float().T
    if b is not None:
        ref = ref + b
    if gelu:
        ref = torch.nn.functional.gelu(ref)
    da, dw = a.cuda(), w.cuda()
    db = b.cuda() if b is not None else None
    if mode == OUT_BF16:
        out = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device="cuda")
        ok(lib().b200x_gemm_bf16(P(da), lda, P(dw), K, M, N, K, bn, P(out), N, mode, P(db), int(gelu), P(None), P(None), 0, 0, 0, 0, P(None)))
        got = out.float().cpu()
        tol = 2e-2
    elif mode == OUT_RESID:
        res = torch.randn(M, N, generator=g)
        ref = ref + res
        out = res.cuda().clone()
        ok(lib().b200x_gemm_bf16(P(da), lda, P(dw), K, M, N, K, bn, P(out), N, mode, P(db), int(gelu), P(out), P(None), 0, 0, 0, 0, P(None)))
        got = out.cpu()
        tol = 2e-4
    else:
        gin, gout, goff = group
        pe = torch.randn(gin, N, generator=g)
        copies = M // gin
        out = torch.zeros(copies * gout, N, device="cuda")
        ok(lib().b200x_gemm_bf16(P(da), lda, P(dw), K, M, N, K, bn, P(out), N, mode, P(db), int(gelu), P(None), P(D(pe)), gin, gout, goff, 0, P(None)))
        full = out.cpu().reshape(copies, gout, N)
        got = full[:, goff:goff + gin].reshape(M, N)
        ref = (ref.reshape(copies, gin, N) + pe).reshape(M, N)
        assert full[:, :goff].abs().max() == 0 if goff else True
        tol = 2e-4
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= tol * max(scale, 1.0), f"max err {err} (scale {scale})"


@pytest.mark.parametrize("bn", [128, 192, 208, 256])
def test_gemm_small_all_tiles(bn):
    _run(300, 416, 384, bn, OUT_BF16, bias=True, gelu=False)


def test_gemm_qkv_shape():
    _run(2 * 1376, 1152, 384, 192, OUT_BF16, bias=False)


def test_gemm_fc1_gelu_padded_hidden():
    _run(1376 + 77, 1040, 384, 208, OUT_BF16, bias=True, gelu=True)


def test_gemm_fc2_residual_k1040():
    _run(1376 + 77, 384, 1040, 192, OUT_RESID, bias=True)


def test_gemm_many_tiles_persistent():
    _run(148 * 128 * 2 + 50, 384, 384, 192, OUT_RESID, bias=True)   # > 1 tile per CTA, exercises both TMEM stages


def test_gemm_token_mode_temporal():
    _run(2 * 1248, 384, 384, 192, OUT_TOKEN, bias=False, gelu=True, group=(1248, 1376, 0))


def test_gemm_token_mode_spectral_k3744():
    _run(2 * 128, 384, 3744, 192, OUT_TOKEN, bias=False, gelu=True, group=(128, 1376, 1248))


def test_gemm_rejects_bad_arguments():
    a = torch.zeros(128, 384, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(RuntimeError):
        ok(lib().b200x_gemm_bf16(P(a), 384, P(a), 384, 128, 100, 384, 192, P(a), 100, 0, P(None), 0, P(None), P(None), 0, 0, 0, 0, P(None)))
    with pytest.raises(RuntimeError):
        ok(lib().b200x_gemm_bf16(P(a), 384, P(a), 384, 128, 128, 384, 64, P(a), 128, 0, P(None), 0, P(None), P(None), 0, 0, 0, 0, P(None)))


@pytest.mark.parametrize("M,N,K,bn,mode,gelu", [
    (148 * 128 * 2 + 50, 384, 384, 192, OUT_RESID, False),     # ragged last row tile: warps without rows skip their stores
    (16 * 1376, 1152, 384, 192, OUT_BF16, False),
    (16 * 1376, 1040, 384, 208, OUT_BF16, True),
    (16 * 1376, 384, 1040, 192, OUT_RESID, False),
])
def test_cta_pair_kernel_is_deterministic_and_direction_independent(M, N, K, bn, mode, gelu):
    """The cta_group::2 kernel (256-row tiles across two SMs) launched repeatedly, forwards and in reverse tile order, on the
    same operands: every launch must give the same bits (shared-memory slab / TMEM stage races show up as sporadic
    mismatching tiles), and those bits must match the fp32 matmul of the bf16-rounded operands."""
    g = torch.Generator(device="cpu").manual_seed(3)
    a = torch.randn(M, K, generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(torch.bfloat16)
    b = torch.randn(N, generator=g)
    res = torch.randn(M, N, generator=g)
    da, dw, db, dres = a.cuda(), w.cuda(), b.cuda(), res.cuda()

    def once(reverse):
        if mode == OUT_RESID:
            out = dres.clone()
            ok(lib().b200x_gemm_bf16(P(da), K, P(dw), K, M, N, K, bn, P(out), N, mode, P(db), 0, P(out), P(None), 0, 0, 0, reverse, P(None)))
        else:
            out = torch.zeros(M, N, dtype=torch.bfloat16, device="cuda")
            ok(lib().b200x_gemm_bf16(P(da), K, P(dw), K, M, N, K, bn, P(out), N, mode, P(db), int(gelu), P(None), P(None), 0, 0, 0, reverse, P(None)))
        torch.cuda.synchronize()
        return out.float()

    first = once(0)
    ref = da.float() @ dw.float().T + db
    if gelu:
        ref = torch.nn.functional.gelu(ref)
    if mode == OUT_RESID:
        ref = ref + dres
    tol = (2e-4 if mode == OUT_RESID else 2e-2) * max(1.0, ref.abs().max().item())
    assert (first - ref).abs().max().item() <= tol
    for i in range(12):
        assert torch.equal(once(i & 1), first)


@pytest.mark.parametrize("M,N,K,bias", [
    (300, 384, 384, True),              # ragged, fewer row tiles than CTA pairs
    (2 * 1376 + 77, 384, 384, True),    # attention projection
    (2 * 1376 + 77, 384, 1040, True),   # fc2: padded hidden width, K not a multiple of 64
    (20 * 1376, 384, 384, False),       # 108 row tiles over 74 pairs: x_ready / ln_done are reused
    (64 * 1376, 384, 1040, True),       # 344 row tiles: 4-5 per CTA pair
    (5 * 1376, 256, 384, True),         # LayerNorm width 256 (two vectors per lane)
])
def test_residual_gemm_with_layernorm_tail_is_bit_identical_to_the_two_kernels(M, N, K, bias):
    """b200x_gemm_resid_ln_bf16 (x += a.W^T + b by TMA reduce-add, then LayerNorm of the CTA's own rows from L2) against
    b200x_gemm_bf16(RESID) followed by b200x_layernorm on the same inputs: x and h carry the same bits, in both traversal
    directions, launch after launch."""
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).to(torch.bfloat16).cuda()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(torch.bfloat16).cuda()
    b = torch.randn(N, generator=g).cuda() if bias else None
    x0 = (torch.randn(M, N, generator=g) * 1.7 + 0.3).cuda()
    gamma = (1.0 + 0.1 * torch.randn(N, generator=g)).cuda()
    beta = (0.1 * torch.randn(N, generator=g)).cuda()
    x_ref = x0.clone()
    ok(lib().b200x_gemm_bf16(P(a), K, P(w), K, M, N, K, 192, P(x_ref), N, OUT_RESID, P(b), 0, P(x_ref), P(None), 0, 0, 0, 0, P(None)))
    h_ref = torch.zeros(M, N, dtype=torch.bfloat16, device="cuda")
    ok(lib().b200x_layernorm(P(x_ref), M, N, P(gamma), P(beta), P(None), P(None), 0, 0, 1e-6, P(h_ref), P(None), 0, P(None)))
    # pin the pair itself against plain torch
    want = x0 + a.float() @ w.float().T + (b if bias else 0.0)
    assert (x_ref - want).abs().max().item() <= 2e-4 * max(1.0, want.abs().max().item())
    hw = torch.nn.functional.layer_norm(x_ref, (N,), gamma, beta, 1e-6)
    assert (h_ref.float() - hw).abs().max().item() <= 2e-2 * max(1.0, hw.abs().max().item())
    for i in range(6):
        x = x0.clone()
        h = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device="cuda")
        ok(lib().b200x_gemm_resid_ln_bf16(P(a), K, P(w), K, M, N, K, P(x), N, P(b), P(gamma), P(beta), 1e-6, P(h), N, i & 1, P(None)))
        torch.cuda.synchronize()
        assert torch.equal(x, x_ref), f"launch {i}: x differs by {(x - x_ref).abs().max().item()}"
        assert torch.equal(h, h_ref), f"launch {i}: h differs by {(h.float() - h_ref.float()).abs().max().item()}"


def test_residual_gemm_with_layernorm_tail_rejects_bad_arguments():
    a = torch.zeros(128, 384, dtype=torch.bfloat16, device="cuda")
    w = torch.zeros(512, 384, dtype=torch.bfloat16, device="cuda")
    x = torch.zeros(128, 512, device="cuda")
    h = torch.zeros(128, 512, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(RuntimeError):       # LayerNorm width 512 > 384: a row tile no longer fits two column tiles of 192
        ok(lib().b200x_gemm_resid_ln_bf16(P(a), 384, P(w), 384, 128, 512, 384, P(x), 512, P(None), P(x), P(x), 1e-6, P(h), 512, 0, P(None)))
    with pytest.raises(RuntimeError):       # NULL LayerNorm parameters
        ok(lib().b200x_gemm_resid_ln_bf16(P(a), 384, P(w), 384, 128, 384, 384, P(x), 512, P(None), P(None), P(None), 1e-6, P(h), 512, 0, P(None)))


@pytest.mark.parametrize("M,N,K,bn,bias,gelu", [
    (300, 1152, 384, 192, False, False),            # one ragged row tile
    (2 * 1376 + 77, 1152, 384, 192, True, False),   # QKV with bias, ragged
    (40 * 1376, 1152, 384, 192, False, False),      # 215 row tiles over 74 pairs: the stationary k-blocks are refilled
    (9 * 1376, 768, 256, 192, True, False),         # four column tiles, four k-blocks
    (9 * 1376 + 5, 1040, 384, 208, True, True),     # fc1 shape: GELU epilogue, 16-column tail group
])
def test_a_stationary_gemm_is_bit_identical_to_the_pair_kernel(M, N, K, bn, bias, gelu):
    """b200x_gemm_bf16_astationary (row tile of A resident in shared memory, rotated column-tile order, deep weight ring) against
    the CTA-pair kernel behind b200x_gemm_bf16 at the same shapes: same bits, both traversal directions, launch after launch."""
    g = torch.Generator(device="cpu").manual_seed(M + N)
    a = torch.randn(M, K, generator=g).to(torch.bfloat16).cuda()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(torch.bfloat16).cuda()
    b = torch.randn(N, generator=g).cuda() if bias else None
    ref = torch.zeros(M, N, dtype=torch.bfloat16, device="cuda")
    ok(lib().b200x_gemm_bf16(P(a), K, P(w), K, min(M, 4 * 73 * 256), N, K, bn, P(ref), N, OUT_BF16, P(b), int(gelu), P(None), P(None), 0, 0, 0, 0, P(None)))
    if M > 4 * 73 * 256:                # keep the reference on the pair kernel (the dispatcher switches at 4 row tiles per pair)
        m0 = 4 * 73 * 256
        ok(lib().b200x_gemm_bf16(P(a[m0:]), K, P(w), K, M - m0, N, K, bn, P(ref[m0:]), N, OUT_BF16, P(b), int(gelu), P(None), P(None), 0, 0, 0, 0, P(None)))
    want = a.float() @ w.float().T + (b if bias else 0.0)
    if gelu:
        want = torch.nn.functional.gelu(want)
    assert (ref.float() - want).abs().max().item() <= 2e-2 * max(1.0, want.abs().max().item())
    for i in range(6):
        out = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device="cuda")
        ok(lib().b200x_gemm_bf16_astationary(P(a), K, P(w), K, M, N, K, bn, P(out), N, P(b), int(gelu), i & 1, P(None)))
        torch.cuda.synchronize()
        assert torch.equal(out, ref), f"launch {i}: {(out.float() - ref.float()).abs().max().item()}"


@pytest.mark.parametrize("batch,K,gelu", [(3, 3744, True), (2, 96, False), (5, 640, True)])
def test_token_gemm_with_m_major_operand_matches_the_k_major_call(batch, K, gelu):
    """b200x_gemm_tokens_mmajor reads A as [batch][K][128] through an M-major tcgen05 descriptor; the regular token-mode call
    on the transposed copy [batch * 128][K] must give the same bits (same products, same accumulation order)."""
    N, T, off = 384, 1376, 1248
    g = torch.Generator(device="cpu").manual_seed(K)
    img = torch.randn(batch, K, 128, generator=g).to(torch.bfloat16).cuda()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(torch.bfloat16).cuda()
    b = torch.randn(N, generator=g).cuda()
    pe = torch.randn(128, N, generator=g).cuda()
    a_k = img.transpose(1, 2).contiguous().reshape(batch * 128, K)
    ref = torch.zeros(batch * T, N, device="cuda")
    ok(lib().b200x_gemm_bf16(P(a_k), K, P(w), K, batch * 128, N, K, 192, P(ref), N, OUT_TOKEN, P(b), int(gelu), P(None), P(pe), 128, T, off, 0, P(None)))
    want = a_k.float() @ w.float().T + b
    if gelu:
        want = torch.nn.functional.gelu(want)
    want = (want.reshape(batch, 128, N) + pe).reshape(batch * 128, N)
    got_ref = ref.reshape(batch, T, N)[:, off:off + 128].reshape(batch * 128, N)
    assert (got_ref - want).abs().max().item() <= 2e-4 * max(1.0, want.abs().max().item())
    for _ in range(3):
        out = torch.zeros(batch * T, N, device="cuda")
        ok(lib().b200x_gemm_tokens_mmajor(P(img), batch, K, P(w), K, N, P(out), N, P(b), int(gelu), P(pe), T, off, P(None)))
        torch.cuda.synchronize()
        assert torch.equal(out, ref), (out - ref).abs().max().item()
