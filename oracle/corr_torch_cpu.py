"""CPU baseline arm: the reference's CorrBlock restated with the SAME ATen ops it calls
(torch.matmul, F.avg_pool2d, F.grid_sample), so timing it on the host cores times the
reference's own CPU path for the hot path.  TEST / BASELINE INFRASTRUCTURE ONLY
(see oracle/__init__.py): bench.py's `cpu_baseline` and `--impl reference` legs and
tests/ may use it; the product never does.

Reference: FF_RAFT_Core/corr.py:12-60, utils/utils.py:57-71.  Checked against the numpy
oracle and the golden vectors in tests/test_oracle_golden.py.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


class TorchCorrBlock:
    def __init__(self, fmap1, fmap2, num_levels: int = 4, radius: int = 4):
        self.num_levels, self.radius = num_levels, radius
        b, d, h, w = fmap1.shape
        a = fmap1.reshape(b, d, h * w).transpose(1, 2)
        vol = torch.matmul(a, fmap2.reshape(b, d, h * w)) / math.sqrt(d)       # corr.py:58-60
        level = vol.reshape(b * h * w, 1, h, w)
        self.corr_pyramid = [level]
        for _ in range(num_levels - 1):                                          # corr.py:24-27
            level = F.avg_pool2d(level, 2, stride=2)
            self.corr_pyramid.append(level)

    def __call__(self, coords):
        r = self.radius
        b, _, h, w = coords.shape
        centre = coords.permute(0, 2, 3, 1).reshape(b * h * w, 1, 1, 2)
        offs = torch.linspace(-r, r, 2 * r + 1, device=coords.device)
        # first window index moves x, second moves y (corr.py:37-43)
        delta = torch.stack(torch.meshgrid(offs, offs, indexing="ij"), dim=-1).view(1, 2 * r + 1, 2 * r + 1, 2)
        outs = []
        for i, level in enumerate(self.corr_pyramid):
            pos = centre / 2 ** i + delta
            hh, ww = level.shape[-2:]
            gx = 2 * pos[..., 0] / (ww - 1) - 1                                 # utils.py:61-62
            gy = 2 * pos[..., 1] / (hh - 1) - 1
            s = F.grid_sample(level, torch.stack([gx, gy], dim=-1), align_corners=True)
            outs.append(s.view(b, h, w, -1))
        return torch.cat(outs, dim=-1).permute(0, 3, 1, 2).contiguous().float()
