"""TEST INFRASTRUCTURE -- runs the REFERENCE's own PWC correlation CUDA kernels (oracle/_ref/libpwc_ref_cuda.so, built
by oracle/build_pwc_ref_cuda.py from core/models/ff-pwcnet/PWCNet_Core/correlation.py) on the current GPU.
Only tests/ and oracle/make_golden_pwc.py may import this module."""
from __future__ import annotations

import ctypes
import os

import torch

_SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libpwc_ref_cuda.so")
_lib = None


def available() -> bool:
    return os.path.exists(_SO)


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_SO)
    return _lib


def shapes():
    L = _load()
    out = []
    for i in range(L.pwc_ref_num_shapes()):
        v = [ctypes.c_int() for _ in range(4)]
        assert L.pwc_ref_shape(i, *[ctypes.byref(x) for x in v]) == 0
        out.append(tuple(x.value for x in v))
    return out


def run(idx: int, one: torch.Tensor, two: torch.Tensor, gout: torch.Tensor, backward: bool = True):
    """-> (out, grad_one, grad_two) as computed by the reference kernels for shape `idx` (inputs must match it);
    backward=False stops after the forward (gradients stay zero)."""
    L = _load()
    b, c, h, w = shapes()[idx]
    assert tuple(one.shape) == tuple(two.shape) == (b, c, h, w) and tuple(gout.shape) == (b, 81, h, w)
    one, two, gout = one.contiguous().float(), two.contiguous().float(), gout.contiguous().float()
    rbot0 = torch.zeros(b, h + 8, w + 8, c, device=one.device)
    rbot1 = torch.zeros_like(rbot0)
    out = torch.zeros(b, 81, h, w, device=one.device)
    gone, gtwo = torch.zeros_like(one), torch.zeros_like(two)
    torch.cuda.synchronize()
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    rc = L.pwc_ref_forward(idx, p(one), p(two), p(rbot0), p(rbot1), p(out))
    assert rc == 0, f"pwc_ref_forward -> {rc}"
    if backward:
        rc = L.pwc_ref_backward(idx, p(rbot0), p(rbot1), p(gout), p(gone), p(gtwo))
        assert rc == 0, f"pwc_ref_backward -> {rc}"
    return out, gone, gtwo
