"""Generate tests/golden/flowformer_ops.npz by running the UNMODIFIED reference FlowFormer functions on CPU.

TEST INFRASTRUCTURE ONLY.  The three FlowFormer operations that reuse the correlation kernels are plain
torch + einops functions inside modules whose import chain needs `timm`, `loguru` and (sic) `turtle`, none of which are
installed: those names are registered as inert stand-in modules so that the reference files import; the functions under
test never touch them.  Reference (core/models/ff-flowformer/FF_FlowFormer_Core/FlowFormer/LatentCostFormer/):
  encoder.py:337-348   MemoryEncoder.corr                 einsum cost volume, [B, heads, h, w, h, w], unscaled
  decoder.py:185-203   MemoryDecoder.encode_flow_token    single-level 9x9 window lookup of per-query cost maps
  decoder.py:119-149   ReverseCostExtractor.forward       cost maps re-sampled at coords1, then 9x9 windows around coords0
"""
from __future__ import annotations

import os
import sys
import warnings
from types import SimpleNamespace
from unittest.mock import MagicMock

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/core/models/ff-flowformer"
GOLD = os.path.join(ROOT, "tests", "golden")


def load_reference():
    for name in ("timm", "timm.data", "timm.models", "timm.models.layers", "timm.models.registry", "timm.models.helpers",
                 "timm.models.vision_transformer", "loguru", "turtle", "tkinter"):
        sys.modules.setdefault(name, MagicMock())
    sys.path.insert(0, REF)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from FF_FlowFormer_Core.FlowFormer.LatentCostFormer import decoder, encoder
    return encoder, decoder


def main():
    enc, dec = load_reference()
    rng = np.random.RandomState(4242)
    out = {}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for tag, (b, dim, h, w, heads) in {"a": (2, 64, 12, 16, 1), "b": (1, 96, 9, 13, 2)}.items():
            f1 = rng.standard_normal((b, dim, h, w)).astype(np.float32)
            f2 = rng.standard_normal((b, dim, h, w)).astype(np.float32)
            cost = enc.MemoryEncoder.corr(SimpleNamespace(cfg=SimpleNamespace(cost_heads_num=heads)), torch.from_numpy(f1), torch.from_numpy(f2))
            cost_maps = cost.permute(0, 2, 3, 1, 4, 5).contiguous().view(b * h * w, heads, h, w)   # encoder.py:361-362 layout
            ys, xs = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
            grid = np.repeat(np.stack([xs, ys], 0).astype(np.float32)[None], b, 0)
            coords0 = grid + rng.standard_normal(grid.shape).astype(np.float32) * 0.6
            coords1 = grid + rng.standard_normal(grid.shape).astype(np.float32) * 2.5
            coords1[0, :, 0, 0] = [-3.0, 2.0]            # re-sampling position outside the map
            coords1[0, :, 1, 1] = [2.0, 3.0]             # exactly on a pixel
            tok = dec.MemoryDecoder.encode_flow_token(None, cost_maps, torch.from_numpy(coords1))
            rev = dec.ReverseCostExtractor(SimpleNamespace()).forward(cost_maps, torch.from_numpy(coords0), torch.from_numpy(coords1))
            out.update({f"{tag}_shape": np.array([b, dim, h, w, heads]), f"{tag}_fmap1": f1, f"{tag}_fmap2": f2,
                        f"{tag}_cost": cost.numpy(), f"{tag}_coords0": coords0, f"{tag}_coords1": coords1,
                        f"{tag}_flow_token": tok.numpy(), f"{tag}_reverse": rev.numpy()})
            print(tag, "cost", tuple(cost.shape), "token", tuple(tok.shape), "reverse", tuple(rev.shape))
    path = os.path.join(GOLD, "flowformer_ops.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) >> 10, "KiB")


if __name__ == "__main__":
    main()
