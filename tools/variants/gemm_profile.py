#!/usr/bin/env python
"""In-kernel issuer cycle counters of the GEMM kernel for the encoder shapes: python tools/gemm_profile.py"""
import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audio_deepfake_explainability_b200 import _lib
lib = _lib.load()
copies, T, D, HP = 16, 1376, 384, 1040
M = copies * T
g = torch.Generator(device="cuda").manual_seed(0)
rnd = lambda *s: (torch.randn(*s, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
P = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)
x = torch.randn(M, D, device="cuda", generator=g)
h, qkv, att, hid = rnd(M, D), rnd(M, 3 * D), rnd(M, D), rnd(M, HP)
w_qkv, w_proj, w_fc1, w_fc2 = rnd(3 * D, D), rnd(D, D), rnd(HP, D), rnd(D, HP)
b_d, b_h = torch.randn(D, device="cuda"), torch.randn(HP, device="cuda")
prof = torch.zeros(148, 8, dtype=torch.int64, device="cuda")
cases = [("qkv  N=1152 K=384", lambda bn: lib.b200x_gemm_bf16(P(h), D, P(w_qkv), D, M, 3 * D, D, bn, P(qkv), 3 * D, 0, P(None), 0, P(None), P(None), 0, 0, 0, P(None))),
         ("proj N=384  K=384", lambda bn: lib.b200x_gemm_bf16(P(att), D, P(w_proj), D, M, D, D, bn, P(x), D, 1, P(b_d), 0, P(x), P(None), 0, 0, 0, P(None))),
         ("fc1  N=1040 K=384", lambda bn: lib.b200x_gemm_bf16(P(h), D, P(w_fc1), D, M, HP, D, bn, P(hid), HP, 0, P(b_h), 1, P(None), P(None), 0, 0, 0, P(None))),
         ("fc2  N=384  K=1040", lambda bn: lib.b200x_gemm_bf16(P(hid), HP, P(w_fc2), HP, M, D, HP, bn, P(x), D, 1, P(b_d), 0, P(x), P(None), 0, 0, 0, P(None)))]
for bres in (0,):
    lib.b200x_debug_gemm_bres(C.c_int(bres))
    for name, fn in cases:
        for bn in (192, 208) if "fc1" in name else (192,):
            lib.b200x_debug_gemm_profile(C.c_void_p(0))
            for _ in range(3): _lib.check(fn(bn))
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20): _lib.check(fn(bn))
            e1.record(); torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 50
            prof.zero_()
            lib.b200x_debug_gemm_profile(C.c_void_p(prof.data_ptr()))
            _lib.check(fn(bn)); torch.cuda.synchronize()
            p = prof.cpu().double()
            p = p[p[:, 3] > 0]
            span = (p[:, 5].max() - p[:, 4].min()) / 1e3
            e = prof.cpu()
            e0, e1, ew = e[0::2, 6].double(), e[0::2, 7].double(), e[1::2, 6].double()
            nz = e0 != 0
            if nz.any():
                rd, ld, cp, st = (e[0::2, 6][nz] >> 32).double().mean(), (e[0::2, 6][nz] & 0xffffffff).double().mean(), (e[0::2, 7][nz] >> 32).double().mean(), (e[0::2, 7][nz] & 0xffffffff).double().mean()
                print(f"   epilogue warp 2 per CTA: wait tfull {ew[nz].mean():8.0f}  wait slab read {rd:8.0f}  tmem ld {ld:8.0f}  compute+st.shared {cp:8.0f}  fence+tma {st:8.0f}")
            print(f"{name} bn={bn}: {us:6.1f} us/launch (events, 20 launches), in-kernel span {span:6.1f} us | issuer per CTA: total {p[:, 2].mean():8.0f} clk (max {p[:, 2].max():8.0f}), wait loads {p[:, 0].mean():8.0f}, "
                  f"wait epilogue {p[:, 1].mean():8.0f}, tiles {p[:, 3].mean():4.1f} (max {p[:, 3].max():.0f}) -> {p[:, 2].mean() / p[:, 3].mean():6.0f} clk/tile; SM clock ~{p[:, 2].max() / max(span, 1e-3) / 1e3:5.2f} GHz x span")
