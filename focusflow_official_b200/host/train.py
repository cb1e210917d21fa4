"""One FocusRAFT training step around the B200 correlation kernels (BASELINE config 5).

Mirrors the reference's loop body, ``core/models/ff-raft/train.py:296-328``, with the ``ffraft_chairs_orb.yaml``
settings every experiment shares: ``model.train()``, ``zero_grad``, 12-iteration forward returning all predictions,
``MixLoss`` (gamma 0.8, lamda 1, 1x1 kernel), ``loss *= world_size`` under DDP (train.py:313-314), backward (stock DDP
bucketed NCCL all-reduce of the 7.66 M parameter gradients overlapping it), ``clip_grad_norm_(1.0)``, ``AdamW`` step,
``OneCycleLR`` step.  ``MIXED_PRECISION`` is false in every config, so there is no autocast / GradScaler work to do.
The correlation path has no parameters: nothing of ours precedes the collective, so there is nothing to fuse with it.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from .losses import build_losses


def parallel_model(model: nn.Module, device: torch.device, rank: int, local_rank: int) -> nn.Module:
    """common.py:45-50: stock DistributedDataParallel when launched by torchrun on a GPU, else the model itself."""
    if device.type != "cpu" and rank != -1:
        from torch.nn.parallel import DistributedDataParallel as DDP

        return DDP(model, device_ids=[local_rank], output_device=local_rank, find_unused_parameters=False)
    return model


class TrainStep:
    """Optimizer, schedule and loss of ``train.py:221-280`` + the loop body ``:296-328`` as a callable."""

    def __init__(self, model: nn.Module, world_size: int = 1, iters: int = 12, lr: float = 4e-4, weight_decay: float = 1e-5,
                 eps: float = 1e-8, num_steps: int = 250000, clip: float = 1.0, loss_type: str = "MixLoss", gamma: float = 0.8,
                 max_flow: float = 400, loss_kernel_size: int = 1, loss_sigma: float = 0.01, lamda: float = 1.0,
                 add_noise: bool = False, freeze_bn: bool = False):
        self.model = model
        self.world_size = world_size
        self.iters = iters
        self.clip = clip
        self.add_noise = add_noise
        self.freeze_bn = freeze_bn
        params = [p for p in model.parameters() if p.requires_grad]
        self.optimizer = torch.optim.AdamW(params, lr=lr, weight_decay=weight_decay, eps=eps)
        self.scheduler = torch.optim.lr_scheduler.OneCycleLR(self.optimizer, lr, num_steps + 100, pct_start=0.05,
                                                             cycle_momentum=False, anneal_strategy="linear")
        self.loss_function = build_losses(loss_type, gamma=gamma, max_flow=max_flow, kernel_size=loss_kernel_size,
                                          sigma=loss_sigma, lamda=lamda)
        self.grad_norm: Optional[torch.Tensor] = None

    def _body(self) -> nn.Module:
        return self.model.module if hasattr(self.model, "module") else self.model

    def __call__(self, image1, image2, flow, mask1, mask2, valid):
        self.model.train()
        if self.freeze_bn:                                   # train.py:192-193: every stage except chairs
            self._body().flow_net.freeze_bn()
        self.optimizer.zero_grad()
        if self.add_noise:                                   # train.py:304-307
            stdv = float(torch.empty(1).uniform_(0.0, 5.0))
            image1 = (image1 + stdv * torch.randn_like(image1)).clamp(0.0, 255.0)
            image2 = (image2 + stdv * torch.randn_like(image2)).clamp(0.0, 255.0)
        preds = self.model(image1, image2, mask1, mask2, raft_iters=self.iters)
        loss, metrics = self.loss_function(preds, flow, valid, mask1)
        if self.world_size > 1:
            loss = loss * self.world_size                    # train.py:313-314
        loss.backward()
        self.grad_norm = torch.nn.utils.clip_grad_norm_(self.model.parameters(), self.clip)
        self.optimizer.step()
        self.scheduler.step()
        return loss.detach(), metrics
