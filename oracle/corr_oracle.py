"""numpy restatement of the RAFT-style correlation block used by FocusRAFT.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Every function states the fp32 arithmetic of the reference step by step, so the
result can be compared with the CUDA kernels at the tolerance BASELINE.json
names (volume 1e-3 rel, pyramid bit-exact given L0, lookup 1e-5).

Reference (all under ``/root/reference/core/models/ff-raft/FF_RAFT_Core/``):
  * ``corr.py:52-60``   CorrBlock.corr          -> :func:`volume`
  * ``corr.py:24-27``   avg_pool2d pyramid      -> :func:`pyramid`
  * ``corr.py:29-50``   CorrBlock.__call__      -> :func:`lookup`
  * ``utils/utils.py:57-71`` bilinear_sampler   -> :func:`_bilinear_taps`
  * ``utils/utils.py:74-77`` coords_grid        -> :func:`coords_grid`
Third-party arithmetic restated here: ATen ``grid_sampler_2d`` (bilinear, zeros
padding, align_corners=True; ``ATen/native/cuda/GridSampler.cuh:22-31`` for the
un-normalisation and the CUDA kernel's nw/ne/sw/se weights) and ATen
``avg_pool2d`` (window sum in (kh, kw) order then one divide).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def coords_grid(batch: int, ht: int, wd: int) -> np.ndarray:
    """``utils.py:74-77``: channel 0 = x (column), channel 1 = y (row)."""
    ys, xs = np.meshgrid(np.arange(ht), np.arange(wd), indexing="ij")
    g = np.stack([xs, ys], axis=0).astype(F32)
    return np.repeat(g[None], batch, axis=0)


def volume(fmap1: np.ndarray, fmap2: np.ndarray) -> np.ndarray:
    """``corr.py:52-60``: corr[b,i,j] = sum_d f1[b,d,i] f2[b,d,j] / sqrt(D).

    fp32 inputs, fp32 accumulate (BLAS sgemm order; the reference's CPU path is
    also an sgemm, its GPU path a TF32 cuBLAS GEMM -- hence the 1e-3 bar).
    Returns ``[B*N, h, w]`` (the ``[B*N, 1, h, w]`` level 0 without the unit dim).
    """
    b, d, h, w = fmap1.shape
    f1 = np.ascontiguousarray(fmap1, dtype=F32).reshape(b, d, h * w)
    f2 = np.ascontiguousarray(fmap2, dtype=F32).reshape(b, d, h * w)
    corr = np.matmul(f1.transpose(0, 2, 1), f2)  # [B, N, N] fp32
    # corr.py:60 divides by sqrt(tensor(dim).float()); a true fp32 division.
    corr = corr / np.sqrt(F32(d))
    return corr.astype(F32).reshape(b * h * w, h, w)


def volume_f64(fmap1: np.ndarray, fmap2: np.ndarray) -> np.ndarray:
    """Exact (fp64) version of :func:`volume`, for precision accounting."""
    b, d, h, w = fmap1.shape
    f1 = fmap1.astype(np.float64).reshape(b, d, h * w)
    f2 = fmap2.astype(np.float64).reshape(b, d, h * w)
    corr = np.matmul(f1.transpose(0, 2, 1), f2) / np.sqrt(np.float64(d))
    return corr.reshape(b * h * w, h, w)


def pool2x2(level: np.ndarray) -> np.ndarray:
    """One ``F.avg_pool2d(corr, 2, stride=2)`` (``corr.py:26``) on ``[Q, h, w]``.

    Odd trailing row/column is dropped (floor); the window is summed in ATen's
    loop order (0,0),(0,1),(1,0),(1,1) and divided once by 4.
    """
    q, h, w = level.shape
    h2, w2 = h // 2, w // 2
    v = level[:, : 2 * h2, : 2 * w2].astype(F32, copy=False)
    a = v[:, 0::2, 0::2]
    b = v[:, 0::2, 1::2]
    c = v[:, 1::2, 0::2]
    d = v[:, 1::2, 1::2]
    s = ((a + b) + c) + d
    return (s / F32(4)).astype(F32)


def pyramid(level0: np.ndarray, num_levels: int = 4) -> list[np.ndarray]:
    """``corr.py:23-27``: [L0, pool(L0), pool(pool(L0)), ...]."""
    out = [np.ascontiguousarray(level0, dtype=F32)]
    for _ in range(num_levels - 1):
        out.append(pool2x2(out[-1]))
    return out


def _axis_taps(centre: np.ndarray, offs: np.ndarray, size: int):
    """Per-tap source index and weights along one axis.

    ``centre`` [Q] fp32 is the level-scaled coordinate, ``offs`` [K] fp32 the
    integer window offsets.  Follows the reference's round trip exactly:
      utils.py:61-62   g  = 2*x/(size-1) - 1            (fp32 mul, div, sub)
      GridSampler.cuh:26  ix = ((g + 1) / 2) * (size-1) (fp32 add, div, mul)
    then the CUDA kernel's corner indices / distances:
      i0 = floor(ix); w0 = (i0 + 1) - ix; w1 = ix - i0.
    Returns (i0 [Q,K] int64, w0 [Q,K] fp32, w1 [Q,K] fp32, finite [Q,K] bool).
    """
    x = (centre[:, None] + offs[None, :]).astype(F32)
    sm1 = F32(size - 1)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        g = (F32(2) * x) / sm1 - F32(1)
        ix = ((g + F32(1)) / F32(2)) * sm1
        f = np.floor(ix)
        w1 = (ix - f).astype(F32)
        w0 = ((f + F32(1)) - ix).astype(F32)
    finite = np.isfinite(ix) & (np.abs(ix) < F32(1e9))
    i0 = np.where(finite, f, -10).astype(np.int64)
    return i0, w0, w1, finite


def lookup_level(level: np.ndarray, cx: np.ndarray, cy: np.ndarray, radius: int) -> np.ndarray:
    """Bilinear (2r+1)^2 window of one level.  ``level`` [Q,h,w]; cx,cy [Q].

    Output ``[Q, (2r+1)^2]`` with k = a*(2r+1)+b <-> (x + a - r, y + b - r):
    ``corr.py:37-43`` stacks meshgrid(dy, dx) and adds it to (x, y), so the
    FIRST window index moves x.  Corner accumulation order and weights are the
    CUDA grid_sampler_2d kernel's: nw, ne, sw, se; each corner dropped
    individually when out of bounds (zeros padding).
    """
    q, h, w = level.shape
    k = 2 * radius + 1
    offs = np.linspace(-radius, radius, k).astype(F32)  # corr.py:37-38
    ix0, wx0, wx1, fx = _axis_taps(cx, offs, w)  # [Q,K] indexed by a
    iy0, wy0, wy1, fy = _axis_taps(cy, offs, h)  # [Q,K] indexed by b

    qi = np.arange(q)[:, None, None]

    def tap(iy, ix):
        ok = (iy >= 0) & (iy < h) & (ix >= 0) & (ix < w)
        v = level[qi, np.clip(iy, 0, h - 1), np.clip(ix, 0, w - 1)]
        return np.where(ok, v, F32(0)).astype(F32)

    X0 = ix0[:, :, None]
    Y0 = iy0[:, None, :]
    X0, Y0 = np.broadcast_arrays(X0, Y0)
    fin = fx[:, :, None] & fy[:, None, :]
    nw = (wx0[:, :, None] * wy0[:, None, :]).astype(F32)
    ne = (wx1[:, :, None] * wy0[:, None, :]).astype(F32)
    sw = (wx0[:, :, None] * wy1[:, None, :]).astype(F32)
    se = (wx1[:, :, None] * wy1[:, None, :]).astype(F32)
    with np.errstate(invalid="ignore", over="ignore"):
        out = tap(Y0, X0) * nw
        out = out + tap(Y0, X0 + 1) * ne
        out = out + tap(Y0 + 1, X0) * sw
        out = out + tap(Y0 + 1, X0 + 1) * se
    out = np.where(fin, out, F32(0)).astype(F32)
    return out.reshape(q, k * k)


def lookup(pyr: list[np.ndarray], coords: np.ndarray, radius: int = 4) -> np.ndarray:
    """``corr.py:29-50``.  ``coords`` [B,2,h,w] (ch0 = x) -> ``[B, L*(2r+1)^2, h, w]``."""
    b, _, h, w = coords.shape
    c = np.ascontiguousarray(coords, dtype=F32).transpose(0, 2, 3, 1).reshape(b * h * w, 2)
    outs = []
    for i, lvl in enumerate(pyr):
        scale = F32(2 ** i)
        cx = (c[:, 0] / scale).astype(F32)  # corr.py:40 true division by 2**i (exact)
        cy = (c[:, 1] / scale).astype(F32)
        outs.append(lookup_level(lvl, cx, cy, radius))
    out = np.concatenate(outs, axis=1).reshape(b, h, w, -1)
    return np.ascontiguousarray(out.transpose(0, 3, 1, 2), dtype=F32)


class CorrBlock:
    """Oracle mirror of ``corr.py:12-60`` (same constructor / call surface)."""

    def __init__(self, fmap1, fmap2, num_levels: int = 4, radius: int = 4):
        self.num_levels = num_levels
        self.radius = radius
        self.corr_pyramid = pyramid(volume(np.asarray(fmap1), np.asarray(fmap2)), num_levels)

    def __call__(self, coords):
        return lookup(self.corr_pyramid, np.asarray(coords), self.radius)


# ----------------------------------------------------------------------------
# gradients: d loss / d fmap through lookup -> pyramid -> volume
# ----------------------------------------------------------------------------

def lookup_level_backward(gout: np.ndarray, shape, cx, cy, radius: int) -> np.ndarray:
    """Adjoint of :func:`lookup_level` w.r.t. ``level`` (coords are detached by
    the caller, ``raft.py:216``).  ``gout`` [Q, (2r+1)^2] -> grad level [Q,h,w]."""
    q, h, w = shape
    k = 2 * radius + 1
    offs = np.linspace(-radius, radius, k).astype(F32)
    ix0, wx0, wx1, fx = _axis_taps(cx, offs, w)
    iy0, wy0, wy1, fy = _axis_taps(cy, offs, h)
    g = gout.reshape(q, k, k).astype(np.float64)
    g = np.where(fx[:, :, None] & fy[:, None, :], g, 0.0)
    grad = np.zeros((q, h, w), dtype=np.float64)
    qi = np.broadcast_to(np.arange(q)[:, None, None], (q, k, k))
    for dy, wy in ((0, wy0), (1, wy1)):
        for dx, wx in ((0, wx0), (1, wx1)):
            X = np.broadcast_to(ix0[:, :, None] + dx, (q, k, k))
            Y = np.broadcast_to(iy0[:, None, :] + dy, (q, k, k))
            ok = (Y >= 0) & (Y < h) & (X >= 0) & (X < w)
            wgt = wx[:, :, None].astype(np.float64) * wy[:, None, :].astype(np.float64)
            np.add.at(grad, (qi[ok], Y[ok], X[ok]), (g * wgt)[ok])
    return grad.astype(F32)


def pool2x2_backward(gl: np.ndarray, shape) -> np.ndarray:
    q, h, w = shape
    out = np.zeros((q, h, w), dtype=F32)
    h2, w2 = h // 2, w // 2
    v = (gl / F32(4)).astype(F32)
    out[:, 0:2 * h2:2, 0:2 * w2:2] = v
    out[:, 0:2 * h2:2, 1:2 * w2:2] = v
    out[:, 1:2 * h2:2, 0:2 * w2:2] = v
    out[:, 1:2 * h2:2, 1:2 * w2:2] = v
    return out
