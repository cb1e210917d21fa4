"""Sequence losses of FocusRAFT training (caller side of the hot path; PyTorch, like the reference keeps them).

Same names, constructor arguments and return values as ``core/models/ff-raft/losses/losses.py``:

  EPELoss(gamma, max_flow)                          losses.py:18-48   RAFT's exponentially weighted L1 sequence loss
  CPCL(gamma, max_flow, kernel_size, sigma)         losses.py:51-93   conditional point control loss: L1 on the key points
  MixLoss(gamma, max_flow, kernel_size, sigma, lamda)  losses.py:96-138  EPELoss + lamda * CPCL   (every shipped config)
  build_losses(loss_type, ...)                      losses/__init__.py:3-11

``forward(flow_preds, flow_gt, valid, mask) -> (loss, metrics)`` with ``metrics = {"epe", "loss"}`` as Python floats.
All three are one weighted sum over the prediction sequence; they share one implementation here.  With the shipped
settings (``LOSS_KERNEL_SIZE: 1``, ``LOSS_SIGMA: 0.01``) the Gaussian that spreads the key-point mask is the 1x1
kernel [1.0].
"""
from __future__ import annotations

import math
from typing import Dict, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


def get_kernel(kernel_size: int, sigma: float) -> torch.Tensor:
    """losses.py:7-15: normalised Gaussian sampled on [-3 sigma, 3 sigma]^2, shape [1, 1, k, k]."""
    ax = torch.linspace(-3.0 * sigma, 3.0 * sigma, kernel_size, dtype=torch.float64)
    g = torch.exp(-(ax[:, None] ** 2 + ax[None, :] ** 2) / (2.0 * sigma ** 2)) / (2.0 * math.pi * sigma ** 2)
    return (g / g.sum()).float().view(1, 1, kernel_size, kernel_size)


class _SequenceLoss(nn.Module):
    """sum_i gamma^(n-i-1) * [ dense * mean(valid |d_i|)  +  point * sum(valid m |d_i|) / sum(m) ]."""

    dense_weight = 1.0
    uses_points = False

    def __init__(self, gamma: float = 0.8, max_flow: float = 400, kernel_size: int = 5, sigma: float = 1.7, lamda: float = 0.8):
        super().__init__()
        self.gamma, self.max_flow = gamma, max_flow
        self.kernel_size, self.sigma, self.lamda = kernel_size, sigma, lamda

    def point_weight(self) -> float:
        return 0.0

    def spread_mask(self, mask: torch.Tensor) -> torch.Tensor:
        m = (mask > 0).float()
        pad = self.kernel_size // 2
        return F.conv2d(F.pad(m, [pad, pad, pad, pad]), get_kernel(self.kernel_size, self.sigma).to(m.device))

    def forward(self, flow_preds: Sequence[torch.Tensor], flow_gt: torch.Tensor, valid: torch.Tensor, mask=None,
                *unused) -> Tuple[torch.Tensor, Dict[str, float]]:
        n = len(flow_preds)
        mag = torch.sum(flow_gt ** 2, dim=1).sqrt()
        ok = (valid >= 0.5) & (mag < self.max_flow)                      # [B, H, W]
        okf = ok[:, None]
        if self.uses_points:
            m = self.spread_mask(mask)
            m_total = m.sum()
        loss = 0.0
        for i, pred in enumerate(flow_preds):
            w = self.gamma ** (n - i - 1)
            err = (pred - flow_gt).abs()
            if self.uses_points:
                loss = loss + self.point_weight() * w * (okf * m * err).sum() / m_total
            if self.dense_weight:
                loss = loss + self.dense_weight * w * (okf * err).mean()
        epe = torch.sum((flow_preds[-1] - flow_gt) ** 2, dim=1).sqrt().view(-1)[ok.view(-1)]
        return loss, {"epe": epe.mean().item(), "loss": loss.detach().item()}


class EPELoss(_SequenceLoss):
    def __init__(self, gamma: float = 0.8, max_flow: float = 400):
        super().__init__(gamma, max_flow)


class CPCL(_SequenceLoss):
    dense_weight = 0.0
    uses_points = True

    def __init__(self, gamma: float = 0.8, max_flow: float = 400, kernel_size: int = 5, sigma: float = 1.7):
        super().__init__(gamma, max_flow, kernel_size, sigma)

    def point_weight(self) -> float:
        return 1.0


class MixLoss(_SequenceLoss):
    uses_points = True

    def point_weight(self) -> float:
        return self.lamda


def build_losses(loss_type: str, gamma=0.8, max_flow=400, kernel_size=5, sigma=1.7, lamda=0.8, **kwargs):
    if loss_type == "EPELoss":
        return EPELoss(gamma, max_flow)
    if loss_type == "CPCL":
        return CPCL(gamma, max_flow, kernel_size, sigma)
    if loss_type == "MixLoss":
        return MixLoss(gamma, max_flow, kernel_size, sigma, lamda)
    raise ValueError(f'"loss_type":"{loss_type}" is not supported.')
