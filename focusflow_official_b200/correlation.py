"""FunctionCorrelation / ModuleCorrelation: drop-in for ``PWCNet_Core/correlation.py:383-395``.

``FunctionCorrelation(tenOne, tenTwo)`` returns the 81-channel 9x9 local cost volume
``[B, 81, H, W]`` (mean over channels, zero outside the image; channel = (dy+4)*9 + (dx+4)),
differentiable w.r.t. both inputs like the reference's ``_FunctionCorrelation``
(``correlation.py:276-380``).  CUDA tensors only: the reference raises ``NotImplementedError``
on CPU (``correlation.py:320-321``) and so does this.  Kernels run on the tensors' device, on torch's CURRENT stream of it
(the reference launches on CuPy's stream, a latent hazard -- SURVEY.md Appendix B.8).
"""
from __future__ import annotations

import torch

from . import _lib

__all__ = ["FunctionCorrelation", "ModuleCorrelation", "correlation_leaky"]


def _launch_fwd(one: torch.Tensor, two: torch.Tensor, leaky: float) -> torch.Tensor:
    b, c, h, w = one.shape
    out = torch.empty((b, 81, h, w), device=one.device, dtype=torch.float32)
    with _lib.on_device(one, two) as stream:
        _lib.check(_lib.lib().ffcorr_pwc81_f32(one.data_ptr(), two.data_ptr(), out.data_ptr(), b, c, h, w, leaky, stream),
                   "ffcorr_pwc81_f32")
    return out


def _prep(one: torch.Tensor, two: torch.Tensor):
    if not (one.is_cuda and two.is_cuda):
        raise NotImplementedError("FunctionCorrelation has no CPU implementation (as in the reference)")
    if one.dim() != 4 or one.shape != two.shape:
        raise ValueError(f"expected two [B, C, H, W] tensors of equal shape, got {tuple(one.shape)} and {tuple(two.shape)}")
    return one.float().contiguous(), two.float().contiguous()


class _FunctionCorrelation(torch.autograd.Function):
    @staticmethod
    def forward(ctx, one, two):
        one, two = _prep(one, two)
        ctx.save_for_backward(one, two)
        return _launch_fwd(one, two, -1.0)

    @staticmethod
    def backward(ctx, grad_output):
        one, two = ctx.saved_tensors
        b, c, h, w = one.shape
        g = grad_output.contiguous().float()
        g1 = torch.empty_like(one) if ctx.needs_input_grad[0] else None
        g2 = torch.empty_like(two) if ctx.needs_input_grad[1] else None
        with _lib.on_device(one, two, g) as stream:
            _lib.check(_lib.lib().ffcorr_pwc81_bwd_f32(one.data_ptr(), two.data_ptr(), g.data_ptr(),
                                                       g1.data_ptr() if g1 is not None else None,
                                                       g2.data_ptr() if g2 is not None else None, b, c, h, w, stream),
                       "ffcorr_pwc81_bwd_f32")
        return g1, g2


def FunctionCorrelation(tenOne, tenTwo):
    return _FunctionCorrelation.apply(tenOne, tenTwo)


def correlation_leaky(tenOne, tenTwo, negative_slope: float = 0.1):
    """Inference-only fusion of ``leaky_relu(FunctionCorrelation(one, two), 0.1)``
    (``ff_pwcnet.py:317,325``) into the kernel epilogue."""
    if torch.is_grad_enabled() and (tenOne.requires_grad or tenTwo.requires_grad):
        return torch.nn.functional.leaky_relu(FunctionCorrelation(tenOne, tenTwo), negative_slope)
    one, two = _prep(tenOne, tenTwo)
    return _launch_fwd(one, two, float(negative_slope))


class ModuleCorrelation(torch.nn.Module):
    def forward(self, tenOne, tenTwo):
        return _FunctionCorrelation.apply(tenOne, tenTwo)
