#!/usr/bin/env python
"""Launch the attention kernel a few times on the engine's chunk shape (for ncu): python tools/attn_only.py [copies] [iters]"""
import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audio_deepfake_explainability_b200 import _lib
lib = _lib.load()
copies = int(sys.argv[1]) if len(sys.argv) > 1 else 16
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
T, H = 1376, 6
g = torch.Generator(device="cuda").manual_seed(0)
qkv = (torch.randn(copies * T, 3 * H * 64, device="cuda", generator=g)).to(torch.bfloat16)
att = torch.zeros(copies * T, H * 64, dtype=torch.bfloat16, device="cuda")
for _ in range(iters):
    _lib.check(lib.b200x_attention(C.c_void_p(qkv.data_ptr()), C.c_void_p(att.data_ptr()), copies, T, H, 64, C.c_void_p(0)))
torch.cuda.synchronize()
print("ok")
