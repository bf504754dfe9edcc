#!/usr/bin/env python
"""SM clock / power while the attention kernel (or a GEMM) runs back to back: python tools/attn_clock.py [copies] [nq] [variant]"""
import ctypes as C, os, subprocess, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audio_deepfake_explainability_b200 import _lib
lib = _lib.load()
copies = int(sys.argv[1]) if len(sys.argv) > 1 else 64
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 1
var = int(sys.argv[3]) if len(sys.argv) > 3 else 256
T, H = 1376, 6
g = torch.Generator(device="cuda").manual_seed(0)
qkv = (torch.randn(copies * T, 3 * H * 64, device="cuda", generator=g)).to(torch.bfloat16)
att = torch.zeros(copies * T, H * 64, dtype=torch.bfloat16, device="cuda")
lib.b200x_debug_attention_tiles_per_cta(C.c_int(nq))
lib.b200x_debug_attention_variant(C.c_int(var))
run = lambda: _lib.check(lib.b200x_attention(C.c_void_p(qkv.data_ptr()), C.c_void_p(att.data_ptr()), copies, T, H, 64, C.c_void_p(0)))
for _ in range(5): run()
torch.cuda.synchronize()
mon = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.active", "--format=csv,noheader", "-lms", "50"], stdout=subprocess.PIPE, text=True)
time.sleep(0.3)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 3000
e0.record()
for _ in range(n): run()
e1.record()
torch.cuda.synchronize()
time.sleep(0.1)
mon.terminate()
out = mon.stdout.read().strip().splitlines()
print(f"nq={nq} variant={var}: {e0.elapsed_time(e1) * 1e3 / n:.1f} us per launch over {n} launches")
print("clock samples:", " | ".join(out[::2][:14]))
