"""FlowFormer's two uses of the same machinery (reference paths relative to
``core/models/ff-flowformer/FF_FlowFormer_Core/FlowFormer/LatentCostFormer``):

* ``cost_volume(fmap1, fmap2)``  <- ``MemoryEncoder.corr`` (``encoder.py:335-347``): the UNSCALED all-pairs volume,
  ``[B, heads, h, w, h, w]``; every shipped config has ``cost_heads_num: 1``.
* ``encode_flow_token(cost_maps, coords)``  <- ``MemoryDecoder.encode_flow_token`` (``decoder.py:185-203``): one-level
  9x9 bilinear window lookup of ``cost_maps [B*h*w, heads, h, w]`` -> ``[B, heads*81, h, w]``; the same tap order and
  ``bilinear_sampler`` as RAFT's ``CorrBlock.__call__``.

Both run the kernels of :mod:`focusflow_official_b200.corr`; inference only.
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
