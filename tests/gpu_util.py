"""Helpers for the GPU parity tests: device buffers come from torch, calls go through the C ABI (ctypes)."""
import ctypes as C

import numpy as np
import torch

from audio_deepfake_explainability_b200 import _lib


def lib():
    return _lib.load()


def P(t):
    """device / host pointer of a torch tensor or numpy array (None -> NULL)."""
    if t is None:
        return C.c_void_p(0)
    if isinstance(t, np.ndarray):
        return C.c_void_p(t.ctypes.data)
    return C.c_void_p(t.data_ptr())


def ok(status, what=""):
    _lib.check(status, what)
    torch.cuda.synchronize()


def bf16_round(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float32)


_KEEP = []


def D(x):
    """Upload a tensor / array to the GPU and keep it alive until the process ends (a temporary passed straight to
    P() would be freed - and its block reused by the next temporary - before the kernel runs)."""
    t = torch.from_numpy(np.ascontiguousarray(x)) if isinstance(x, np.ndarray) else x
    t = t.cuda()
    _KEEP.append(t)
    if len(_KEEP) > 64:
        torch.cuda.synchronize()
        del _KEEP[:32]
    return t
