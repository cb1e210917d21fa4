"""FlowFormer's two uses of the same machinery (reference paths relative to
``core/models/ff-flowformer/FF_FlowFormer_Core/FlowFormer/LatentCostFormer``):

* ``cost_volume(fmap1, fmap2)``  <- ``MemoryEncoder.corr`` (``encoder.py:335-347``): the UNSCALED all-pairs volume,
  ``[B, heads, h, w, h, w]``; every shipped config has ``cost_heads_num: 1``.
* ``encode_flow_token(cost_maps, coords)``  <- ``MemoryDecoder.encode_flow_token`` (``decoder.py:185-203``): one-level
  9x9 bilinear window lookup of ``cost_maps [B*h*w, heads, h, w]`` -> ``[B, heads*81, h, w]``; the same tap order and
  ``bilinear_sampler`` as RAFT's ``CorrBlock.__call__``.

* ``reverse_cost_tokens(cost_maps, coords0, coords1)``  <- ``ReverseCostExtractor.forward`` (``decoder.py:119-149``):
  every query's cost map is re-sampled at ``coords1`` (one bilinear sample per target position, which turns the maps
  "query -> targets" into maps "position -> queries"), then a 9x9 window around ``coords0`` is read from those.  The
  first step is a bilinear blend of four ROWS of the transposed volume, so the whole operation is four window lookups
  on the transposed cost maps (the map index comes from the four taps of ``coords1``) blended with the tap weights.

All run the kernels of :mod:`focusflow_official_b200.corr`; inference only.  Pinned by ``tests/golden/flowformer_ops.npz``:
outputs of the reference's own functions, recorded by the FlowFormer golden-vector generator of the test infrastructure.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from .corr import _lookup_raw, _precision_code, _prep, _require_cuda


def cost_volume(fmap1: torch.Tensor, fmap2: torch.Tensor, heads: int = 1, precision: Optional[str] = None) -> torch.Tensor:
    fmap1, fmap2 = _prep(fmap1, "fmap1"), _prep(fmap2, "fmap2")
    if fmap1.shape != fmap2.shape:
        raise ValueError(f"fmap shapes differ: {tuple(fmap1.shape)} vs {tuple(fmap2.shape)}")
    b, dim, h, w = fmap1.shape
    if heads < 1 or dim % heads:
        raise ValueError(f"{dim} channels do not split into {heads} heads")
    d = dim // heads
    n = h * w
    code = _precision_code(precision)
    L = _lib.lib()
    # heads are independent contractions over their own channel slice: fold them into the batch
    f1 = fmap1.view(b * heads, d, h, w)
    f2 = fmap2.view(b * heads, d, h, w)
    out = torch.empty((b * heads, n, n), device=fmap1.device, dtype=torch.float32)
    ws_bytes = L.ffcorr_volume_workspace_bytes(b * heads, d, h, w, code)
    ws = torch.empty(max(ws_bytes, 1), device=fmap1.device, dtype=torch.uint8)
    with _lib.on_device(fmap1, fmap2) as stream:
        _lib.check(L.ffcorr_volume_scaled_f32(f1.data_ptr(), f2.data_ptr(), out.data_ptr(), b * heads, d, h, w, code, 1.0,
                                              ws.data_ptr() if ws_bytes else None, ws_bytes, stream), "ffcorr_volume_scaled_f32")
    return out.view(b, heads, h, w, h, w)


def encode_flow_token(cost_maps: torch.Tensor, coords: torch.Tensor, radius: int = 4) -> torch.Tensor:
    _require_cuda(cost_maps, "cost_maps")
    _require_cuda(coords, "coords")
    b, two, h1, w1 = coords.shape
    q, heads, h2, w2 = cost_maps.shape
    if two != 2 or q != b * h1 * w1:
        raise ValueError(f"cost_maps {tuple(cost_maps.shape)} / coords {tuple(coords.shape)} do not match")
    if (h2, w2) != (h1, w1):
        raise ValueError("the lookup kernel indexes one map per query pixel of the same size (decoder.py:185-203 use)")
    coords = coords.detach().float().contiguous()
    k2 = (2 * radius + 1) ** 2
    if heads == 1:
        lv = [cost_maps.detach().float().contiguous()]
        return _lookup_raw(lv, _lib.ptr_array(lv), coords, radius)
    outs = []
    for j in range(heads):      # channel order of the reference: head-major, then the 81 taps
        lv = [cost_maps[:, j:j + 1].detach().float().contiguous()]
        outs.append(_lookup_raw(lv, _lib.ptr_array(lv), coords, radius))
    return torch.cat(outs, dim=1).view(b, heads * k2, h1, w1)


def _align_corners_taps(coord: torch.Tensor, size: int):
    """Source index and corner weights of ``bilinear_sampler`` (``utils.py``: x -> 2x/(size-1) - 1 -> grid_sample with
    align_corners=True un-normalising ((g+1)/2)*(size-1)), fp32 step by step like ATen's CUDA kernel."""
    sm1 = float(size - 1)
    g = 2 * coord / sm1 - 1
    ix = ((g + 1) / 2) * sm1
    i0 = torch.floor(ix)
    return i0.long(), (i0 + 1) - ix, ix - i0


def reverse_cost_tokens(cost_maps: torch.Tensor, coords0: torch.Tensor, coords1: torch.Tensor, radius: int = 4) -> torch.Tensor:
    """``ReverseCostExtractor.forward`` (``decoder.py:119-149``): ``[B, heads*81, h, w]``."""
    _require_cuda(cost_maps, "cost_maps")
    b, two, h, w = coords1.shape
    q, heads, h2, w2 = cost_maps.shape
    if two != 2 or coords0.shape != coords1.shape or q != b * h * w or (h2, w2) != (h, w):
        raise ValueError(f"cost_maps {tuple(cost_maps.shape)} / coords {tuple(coords0.shape)}, {tuple(coords1.shape)} do not match")
    n = h * w
    k2 = (2 * radius + 1) ** 2
    # maps "target position -> queries": the volume transposed per batch item and head, [B*N(j), heads, h, w] over p
    ct = cost_maps.detach().float().view(b, n, heads, n).permute(0, 3, 2, 1).contiguous().view(b * n, heads, h, w)
    c1 = coords1.detach().float()
    ix0, wx0, wx1 = _align_corners_taps(c1[:, 0], w)          # [B, h, w]
    iy0, wy0, wy1 = _align_corners_taps(c1[:, 1], h)
    base = (torch.arange(b, device=ct.device) * n).view(b, 1, 1)
    c0 = coords0.detach().float().contiguous()
    out = None
    for dy, wy in ((0, wy0), (1, wy1)):
        for dx, wx in ((0, wx0), (1, wx1)):
            jx, jy = ix0 + dx, iy0 + dy
            ok = (jx >= 0) & (jx < w) & (jy >= 0) & (jy < h)            # zeros padding: the tap contributes nothing outside
            idx = (base + jy.clamp(0, h - 1) * w + jx.clamp(0, w - 1)).view(-1)
            maps = ct.index_select(0, idx)                              # the re-sampled row of the transposed volume
            tok = encode_flow_token(maps, c0, radius)                   # [B, heads*81, h, w]
            wgt = (wx * wy * ok).view(b, 1, h, w)
            out = tok * wgt if out is None else out + tok * wgt
    return out.view(b, heads * k2, h, w)
