#!/usr/bin/env python
"""Per-kernel micro-benchmark (CUDA events, L2-resident chunk shapes as in the engine): python tools/kernel_bench.py [copies]"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audio_deepfake_explainability_b200 import _lib  # noqa: E402

lib = _lib.load()
copies = int(sys.argv[1]) if len(sys.argv) > 1 else 16
T, D, H, HP = 1376, 384, 6, 1040
M = copies * T
P = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)


ITERS = int(os.environ.get("KB_ITERS", "0"))
WARMUP = int(os.environ.get("KB_WARMUP", "3"))


def timeit(fn, flops=None, bytes_=None, name="", iters=20):
    iters = ITERS or iters
    for _ in range(WARMUP):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    extra = ""
    if flops:
        extra += f"  {flops / us / 1e6:8.1f} TFLOP/s"
    if bytes_:
        extra += f"  {bytes_ / us / 1e3:8.1f} GB/s"
    print(f"{name:34s} {us:9.1f} us{extra}", flush=True)
    return us


def rnd(*shape, dtype=torch.bfloat16, scale=1.0):
    return (torch.randn(*shape, device=dev, generator=g) * scale).to(dtype)


x = torch.randn(M, D, device=dev, generator=g)
h = rnd(M, D)
qkv = rnd(M, 3 * D)
att = rnd(M, D)
hid = rnd(M, HP)
w_qkv, w_proj, w_fc1, w_fc2 = rnd(3 * D, D, scale=0.05), rnd(D, D, scale=0.05), rnd(HP, D, scale=0.05), rnd(D, HP, scale=0.05)
b_d, b_h = torch.randn(D, device=dev), torch.randn(HP, device=dev)
gam, bet = torch.ones(D, device=dev), torch.zeros(D, device=dev)
total = 0.0


def gemm(a, lda, w, ldw, m, n, k, bn, out, ldc, mode, bias, gelu, resid):
    _lib.check(lib.b200x_gemm_bf16(P(a), lda, P(w), ldw, m, n, k, bn, P(out), ldc, mode, P(bias), gelu, P(resid), P(None), 0, 0, 0, 0, P(None)))


print(f"copies={copies}  M={M}")
total += timeit(lambda: _lib.check(lib.b200x_layernorm(P(x), M, D, P(gam), P(bet), P(None), P(None), 0, 0, 1e-5, P(h), P(None), 0, P(None))),
                bytes_=M * D * 6, name="layernorm fp32->bf16") * 2
total += timeit(lambda: gemm(h, D, w_qkv, D, M, 3 * D, D, 192, qkv, 3 * D, 0, None, 0, None), flops=2 * M * D * 3 * D, name="gemm qkv   N=1152 K=384  bf16")
total += timeit(lambda: _lib.check(lib.b200x_attention(P(qkv), P(att), copies, T, H, 64, 0, P(None))), flops=4 * copies * H * T * T * 64, name="attention")
total += timeit(lambda: gemm(att, D, w_proj, D, M, D, D, 192, x, D, 1, b_d, 0, x), flops=2 * M * D * D, name="gemm proj  N=384  K=384  resid")
total += timeit(lambda: gemm(h, D, w_fc1, D, M, HP, D, 208, hid, HP, 0, b_h, 1, None), flops=2 * M * D * HP, name="gemm fc1   N=1040 K=384  gelu")
total += timeit(lambda: gemm(hid, HP, w_fc2, HP, M, D, HP, 192, x, D, 1, b_d, 0, x), flops=2 * M * D * HP, name="gemm fc2   N=384  K=1040 resid")
timeit(lambda: _lib.check(lib.b200x_gemm_resid_ln_bf16(P(att), D, P(w_proj), D, M, D, D, P(x), D, P(b_d), P(gam), P(bet), 1e-5, P(h), D, 0, P(None))),
       flops=2 * M * D * D, name="gemm proj + LayerNorm tail")
timeit(lambda: _lib.check(lib.b200x_gemm_resid_ln_bf16(P(hid), HP, P(w_fc2), HP, M, D, HP, P(x), D, P(b_d), P(gam), P(bet), 1e-5, P(h), D, 0, P(None))),
       flops=2 * M * D * HP, name="gemm fc2 + LayerNorm tail")
timeit(lambda: _lib.check(lib.b200x_gemm_bf16_astationary(P(h), D, P(w_qkv), D, M, 3 * D, D, 192, P(qkv), 3 * D, P(None), 0, 0, P(None))),
       flops=2 * M * D * 3 * D, name="gemm qkv A-stationary")
timeit(lambda: _lib.check(lib.b200x_gemm_bf16_astationary(P(h), D, P(w_fc1), D, M, HP, D, 208, P(hid), HP, P(b_h), 1, 0, P(None))),
       flops=2 * M * D * HP, name="gemm fc1 A-stationary (gelu)")
layer_flops = 2 * M * D * (3 * D + D + 2 * 1025) + 4 * copies * H * T * T * 64
print(f"{'one encoder layer (sum)':34s} {total:9.1f} us  {layer_flops / total / 1e6:8.1f} TFLOP/s  -> {12 * total / copies:7.1f} us/eval for 12 layers")

# DSP stage
L = 1920000
n_frames = 1 + L // 512
wave = torch.randn(L, device=dev, generator=g) * 0.1
S = torch.zeros(n_frames, 1028, 2, device=dev)
_lib.check(lib.b200x_stft(P(wave), L, 2048, 512, 0, P(S), 1028, P(None)))
y = torch.zeros(copies, L + 8, device=dev)
wins = torch.tensor([[1024, 2048, 100, 151]] * copies, dtype=torch.int32, device=dev)
timeit(lambda: _lib.check(lib.b200x_stft(P(wave), L, 2048, 512, 0, P(S), 1028, P(None))), name="stft (1 track)", bytes_=L * 4 + n_frames * 1025 * 8)
t_i = timeit(lambda: _lib.check(lib.b200x_istft_masked(P(S), 1028, n_frames, copies, 1, P(wins), 0.0, P(None), P(y), L + 8, P(None), P(None), 0, P(None))),
             name=f"istft_masked x{copies}", bytes_=copies * (n_frames * 1025 * 8 + L * 4), iters=5)
n_cta = -(-n_frames // lib.b200x_mel_frames_per_cta())
db = torch.zeros(copies, n_frames, 128, device=dev)
cmax = torch.zeros(copies, n_cta, device=dev)
t_m = timeit(lambda: _lib.check(lib.b200x_mel_db(P(y), L + 8, L, copies, 16000, 128, 20.0, 8000.0, 1e-10, P(None), 0.0, L, P(db), n_frames, P(cmax), P(None), 0, P(None))),
             name=f"mel_db x{copies}", bytes_=copies * (L * 4 + n_frames * 128 * 4), iters=5)
img_t = torch.zeros(copies, 3744, 128, dtype=torch.bfloat16, device=dev)
img_f = torch.zeros(copies, 128, 3744, dtype=torch.bfloat16, device=dev)
part = torch.zeros(copies * 64, dtype=torch.float64, device=dev)
fl = torch.zeros(copies, device=dev)
t_r = timeit(lambda: _lib.check(lib.b200x_mel_normalize_resize(P(db), n_frames, P(cmax), n_cta, copies, n_frames, 128, 80.0, 1, 1e-6, 3744, P(None), P(None), P(None), P(None), P(part), P(fl), P(img_t), P(img_f), 3744, P(None))),
             name=f"normalize+resize x{copies}", iters=5)
rng = torch.zeros(copies, 2, dtype=torch.int32, device=dev)
_lib.check(lib.b200x_frame_ranges(P(wins), copies, n_frames, P(rng), P(None)))
t_is = timeit(lambda: _lib.check(lib.b200x_istft_masked(P(S), 1028, n_frames, copies, 1, P(wins), 0.0, P(None), P(y), L + 8, P(None), P(rng), 1032, P(None))),
              name=f"istft_masked sparse x{copies}", iters=5)
t_ms = timeit(lambda: _lib.check(lib.b200x_mel_db(P(y), L + 8, L, copies, 16000, 128, 20.0, 8000.0, 1e-10, P(None), 0.0, L, P(db), n_frames, P(cmax), P(rng), 1032, P(None))),
              name=f"mel_db sparse x{copies}", iters=5)
print(f"DSP sparse per eval: {(t_is + t_ms + t_r) / copies:7.1f} us")
print(f"DSP per eval: {(t_i + t_m + t_r) / copies:7.1f} us   transformer per eval: {12 * total / copies:7.1f} us")
