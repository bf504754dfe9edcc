#!/usr/bin/env python
"""In-kernel cycle counters of the attention kernel (diagnostic variant 32): python tools/attn_profile.py [copies]"""
import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audio_deepfake_explainability_b200 import _lib
lib = _lib.load()
copies = int(sys.argv[1]) if len(sys.argv) > 1 else 16
T, H = 1376, 6
g = torch.Generator(device="cuda").manual_seed(0)
qkv = (torch.randn(copies * T, 3 * H * 64, device="cuda", generator=g)).to(torch.bfloat16)
att = torch.zeros(copies * T, H * 64, dtype=torch.bfloat16, device="cuda")
n_cta = 6 * H * copies
prof_all = torch.zeros(n_cta * 44 + 11 * 128, dtype=torch.int64, device="cuda")
prof = prof_all[: n_cta * 44].view(n_cta, 11, 4)
lib.b200x_debug_attention_profile(C.c_void_p(prof_all.data_ptr()))
lib.b200x_debug_attention_tiles_per_cta(C.c_int(2))
lib.b200x_debug_attention_variant(C.c_int(int(sys.argv[2]) if len(sys.argv) > 2 else 32))
for _ in range(3):
    _lib.check(lib.b200x_attention(C.c_void_p(qkv.data_ptr()), C.c_void_p(att.data_ptr()), copies, T, H, 64, C.c_void_p(0)))
torch.cuda.synchronize()
p = prof.cpu().double()
full = p.view(copies * H, 6, 11, 4)[:, :5]          # CTAs with two valid query tiles
iss = full[:, :, 9]
print(f"issuer : wait kv {iss[..., 0].mean():9.0f}  wait p_ready {iss[..., 1].mean():9.0f}  total {iss[..., 2].mean():9.0f} cycles per CTA (11 kv steps)")
sm = full[:, :, 0:8]
print(f"softmax: wait S {sm[..., 0].mean():9.0f}  pass {sm[..., 1].mean():9.0f}  tail {sm[..., 2].mean():9.0f} cycles per CTA;"
      f" per kv step: wait {sm[..., 0].mean() / 11:7.0f} pass {sm[..., 1].mean() / 11:7.0f} tail {sm[..., 2].mean() / 11:7.0f}")
packed = prof.cpu().view(copies * H, 6, 11, 4)[:, :5, 0:8, 3]
ld, mx, pv = (packed >> 42).double().mean() / 11, ((packed >> 21) & 0x1FFFFF).double().mean() / 11, (packed & 0x1FFFFF).double().mean() / 11
print(f"  inside pass per kv step: tmem load+wait {ld:6.0f}  s_free arrive + row max {mx:6.0f}  wait pv_done {pv:6.0f}  (rest = rescale check + exp + P stores)")
for w in range(8):
    print(f"  warp {w}: wait {sm[:, :, w, 0].mean() / 11:7.0f} pass {sm[:, :, w, 1].mean() / 11:7.0f} tail {sm[:, :, w, 2].mean() / 11:7.0f}")
last = p.view(copies * H, 6, 11, 4)[:, 5]           # CTAs with ONE valid query tile (group A alone on the MUFU)
sm1 = last[:, 0:4]
print(f"single-tile CTAs: softmax wait {sm1[..., 0].mean() / 11:7.0f} pass {sm1[..., 1].mean() / 11:7.0f} tail {sm1[..., 2].mean() / 11:7.0f} per kv step;"
      f" issuer total {last[:, 9, 2].mean():9.0f}")

tr = prof_all[n_cta * 44:].cpu().view(11, 128)
ev = []
names = {8: "tma", 9: "mmaA", 10: "mmaB", 0: "smA0", 4: "smB0"}
for w, nm in names.items():
    for v in tr[w].tolist():
        if v == 0: continue
        ev.append((v & 0xFFFFFFFFFFFF, nm, (v >> 56) & 0xFF, (v >> 48) & 0xFF))
ev.sort()
t0 = ev[0][0]
lab = {"tma": {0: "start", 1: "load kv"}, "mmaA": {1: "kv ready", 2: "s_free -> issue S(j+1)", 3: "p_ready -> issue PV(j)"},
       "smA0": {1: "S ready", 2: "row loaded", 3: "max done", 4: "pv_done ok, first P store", 5: "exp done", 6: "p_ready arrived", 7: "  probe: S(j+1) NOT ready", 8: "  probe: S(j+1) ready", 9: "  probe: S(j+1) NOT ready", 10: "  probe: S(j+1) ready"}}
lab["mmaB"] = lab["mmaA"]; lab["smB0"] = lab["smA0"]
print("timeline of CTA (1,0,0): cycles since start")
for t, nm, e, j in ev:
    if j <= 4: print(f"{t - t0:8d}  {nm:5s} j={j:2d}  {lab[nm].get(e, e)}")
