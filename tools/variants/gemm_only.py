#!/usr/bin/env python
"""Launch one encoder GEMM shape a few times (for ncu): python tools/gemm_only.py qkv|proj|fc1|fc2 [iters]"""
import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audio_deepfake_explainability_b200 import _lib
lib = _lib.load()
which = sys.argv[1] if len(sys.argv) > 1 else "fc1"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
copies, T, D, HP = 16, 1376, 384, 1040
M = copies * T
g = torch.Generator(device="cuda").manual_seed(0)
rnd = lambda *s: (torch.randn(*s, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
P = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)
x = torch.randn(M, D, device="cuda", generator=g)
h, qkv, att, hid = rnd(M, D), rnd(M, 3 * D), rnd(M, D), rnd(M, HP)
w_qkv, w_proj, w_fc1, w_fc2 = rnd(3 * D, D), rnd(D, D), rnd(HP, D), rnd(D, HP)
b_d, b_h = torch.randn(D, device="cuda"), torch.randn(HP, device="cuda")
fn = {"qkv": lambda: lib.b200x_gemm_bf16(P(h), D, P(w_qkv), D, M, 3 * D, D, 192, P(qkv), 3 * D, 0, P(None), 0, P(None), P(None), 0, 0, 0, P(None)),
      "proj": lambda: lib.b200x_gemm_bf16(P(att), D, P(w_proj), D, M, D, D, 192, P(x), D, 1, P(b_d), 0, P(x), P(None), 0, 0, 0, P(None)),
      "fc1": lambda: lib.b200x_gemm_bf16(P(h), D, P(w_fc1), D, M, HP, D, 208, P(hid), HP, 0, P(b_h), 1, P(None), P(None), 0, 0, 0, P(None)),
      "fc2": lambda: lib.b200x_gemm_bf16(P(hid), HP, P(w_fc2), HP, M, D, HP, 192, P(x), D, 1, P(b_d), 0, P(x), P(None), 0, 0, 0, P(None))}[which]
for _ in range(iters):
    _lib.check(fn())
torch.cuda.synchronize()
print("ok")
