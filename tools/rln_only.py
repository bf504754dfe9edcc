"""Launch the residual GEMM + LayerNorm tail a few times (ncu target).  usage: python tools/rln_only.py [copies] [K]"""
import sys
import torch
sys.path.insert(0, ".")
from audio_deepfake_explainability_b200 import _lib
lib = _lib.load()
P = lambda t: None if t is None else t.data_ptr()
copies = int(sys.argv[1]) if len(sys.argv) > 1 else 229
K = int(sys.argv[2]) if len(sys.argv) > 2 else 384
M, D = copies * 1376, 384
g = torch.Generator(device="cuda").manual_seed(0)
a = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
w = (torch.randn(D, K, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
x = torch.randn(M, D, device="cuda", generator=g)
h = torch.zeros(M, D, dtype=torch.bfloat16, device="cuda")
b, gam, bet = torch.randn(D, device="cuda"), torch.ones(D, device="cuda"), torch.zeros(D, device="cuda")
for _ in range(4):
    _lib.check(lib.b200x_gemm_resid_ln_bf16(P(a), K, P(w), K, M, D, K, P(x), D, P(b), P(gam), P(bet), 1e-5, P(h), D, 0, None), "rln")
torch.cuda.synchronize()
print("ok", float(h.float().abs().mean()))
