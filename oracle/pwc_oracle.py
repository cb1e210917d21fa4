"""Oracle for the PWC 9x9 local cost volume (TEST INFRASTRUCTURE ONLY).

Two independent statements of ``PWCNet_Core/correlation.py``:
  * :func:`forward_c` -- ctypes binding of ``oracle/pwc_ref.c``, a literal C
    restatement of the CUDA-C strings (``correlation.py:7-102``) including the
    32-lane partial-sum order; also the timed CPU baseline of bench.py.
  * :func:`forward_np` -- the closed form
    ``out[:, (dy+4)*9+(dx+4)] = mean_c(one * shift(two, dy, dx))`` in numpy
    (float64 accumulate), used to cross-check the C file.
Backward (``correlation.py:104-232, 331-380``): :func:`backward_c`, :func:`backward_np`.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libpwc_ref.so")
    src = os.path.join(_HERE, "pwc_ref.c")
    if force or not os.path.exists(so) or (
        os.path.exists(src) and os.path.getmtime(so) < os.path.getmtime(src)
    ):
        subprocess.check_call(["make", "-C", _HERE, "libpwc_ref.so"], stdout=subprocess.DEVNULL)
    return so


def _lib():
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(build())
        fp = ctypes.POINTER(ctypes.c_float)
        for name in ("pwc_ref_forward", "pwc_ref_backward_one", "pwc_ref_backward_two"):
            fn = getattr(lib, name)
            fn.argtypes = [fp, fp, fp] + [ctypes.c_int] * 4
            fn.restype = ctypes.c_int
        _LIB = lib
    return _LIB


def _p(a: np.ndarray):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def forward_c(one: np.ndarray, two: np.ndarray) -> np.ndarray:
    one = np.ascontiguousarray(one, dtype=np.float32)
    two = np.ascontiguousarray(two, dtype=np.float32)
    b, c, h, w = one.shape
    out = np.empty((b, 81, h, w), dtype=np.float32)
    rc = _lib().pwc_ref_forward(_p(one), _p(two), _p(out), b, c, h, w)
    if rc != 0:
        raise MemoryError("pwc_ref_forward")
    return out


def forward_np(one: np.ndarray, two: np.ndarray) -> np.ndarray:
    one = np.asarray(one, dtype=np.float64)
    b, c, h, w = one.shape
    tp = np.zeros((b, c, h + 8, w + 8), dtype=np.float64)
    tp[:, :, 4:4 + h, 4:4 + w] = two
    out = np.empty((b, 81, h, w), dtype=np.float64)
    for dy in range(-4, 5):
        for dx in range(-4, 5):
            sh = tp[:, :, 4 + dy:4 + dy + h, 4 + dx:4 + dx + w]
            out[:, (dy + 4) * 9 + (dx + 4)] = (one * sh).mean(axis=1)
    return out.astype(np.float32)


def backward_c(one: np.ndarray, two: np.ndarray, gout: np.ndarray):
    one = np.ascontiguousarray(one, dtype=np.float32)
    two = np.ascontiguousarray(two, dtype=np.float32)
    gout = np.ascontiguousarray(gout, dtype=np.float32)
    b, c, h, w = one.shape
    g1 = np.empty_like(one)
    g2 = np.empty_like(one)
    _lib().pwc_ref_backward_one(_p(two), _p(gout), _p(g1), b, c, h, w)
    _lib().pwc_ref_backward_two(_p(one), _p(gout), _p(g2), b, c, h, w)
    return g1, g2


def backward_np(one: np.ndarray, two: np.ndarray, gout: np.ndarray):
    one = np.asarray(one, dtype=np.float64)
    two = np.asarray(two, dtype=np.float64)
    g = np.asarray(gout, dtype=np.float64)
    b, c, h, w = one.shape
    tp = np.zeros((b, c, h + 8, w + 8))
    tp[:, :, 4:4 + h, 4:4 + w] = two
    g1 = np.zeros_like(one)
    g2p = np.zeros((b, c, h + 8, w + 8))
    for dy in range(-4, 5):
        for dx in range(-4, 5):
            gk = g[:, (dy + 4) * 9 + (dx + 4)][:, None]  # [B,1,H,W]
            g1 += gk * tp[:, :, 4 + dy:4 + dy + h, 4 + dx:4 + dx + w]
            g2p[:, :, 4 + dy:4 + dy + h, 4 + dx:4 + dx + w] += gk * one
    return (g1 / c).astype(np.float32), (g2p[:, :, 4:4 + h, 4:4 + w] / c).astype(np.float32)
