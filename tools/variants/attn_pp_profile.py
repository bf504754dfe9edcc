#!/usr/bin/env python
"""In-kernel cycle counters of the ping-pong attention kernel: python tools/attn_pp_profile.py [copies] [split]"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audio_deepfake_explainability_b200 import _lib  # noqa: E402

lib = _lib.load()
copies = int(sys.argv[1]) if len(sys.argv) > 1 else 64
split = int(sys.argv[2]) if len(sys.argv) > 2 else 1
T, H = 1376, 6
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(copies * T, 3 * H * 64, device="cuda", generator=g).to(torch.bfloat16)
out = torch.zeros(copies * T, H * 64, dtype=torch.bfloat16, device="cuda")
prof = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
lib.b200x_debug_attention_tiles_per_cta(C.c_int(4 + split))
P = lambda t: C.c_void_p(t.data_ptr())
for _ in range(2):
    _lib.check(lib.b200x_attention(P(qkv), P(out), copies, T, H, 64, C.c_void_p(0)))
lib.b200x_debug_attention_profile(P(prof))
_lib.check(lib.b200x_attention(P(qkv), P(out), copies, T, H, 64, C.c_void_p(0)))
torch.cuda.synchronize()
lib.b200x_debug_attention_profile(C.c_void_p(0))
p = prof.cpu().view(148, 16).double()
n = p[:, 6].clamp(min=1)
print(f"copies={copies} split={split}: softmax warp 0, cycles per pass (mean over CTAs): wait S {float((p[:,0]/n).mean()):.0f}  "
      f"first-tile max pass {float((p[:,1]/n).mean()):.0f}  chunk loop {float((p[:,2]/n).mean()):.0f} (of which wait pv_done {float((p[:,3]/n).mean()):.0f})  "
      f"tail {float((p[:,4]/n).mean()):.0f}  total/pass {float((p[:,5]/n).mean()):.0f}  passes {float(n.mean()):.0f}")
steps = p[:, 11].clamp(min=1) * 2
print(f"issuer per half step: wait p_ready {float((p[:,8]/steps).mean()):.0f}  wait kv/q {float((p[:,9]/steps).mean()):.0f}  total {float((p[:,10]/steps).mean()):.0f}")
