"""FF-PWC host model (PyTorch) around the B200 local-cost-volume kernels (BASELINE config 3).

Re-written caller of ``FunctionCorrelation``; mirrors, attribute name by attribute name, the configuration every
reference experiment uses (``FUSION: parallel``, ``FUSION_TYPE: 1x1conv``, ``MASK_MODAL: point``) so that reference
``ffpwc_*.pth`` state dicts load unchanged:

  FocusPWC           <- PWCNet_Core/ff_pwcnet.py:113-434   (FF_PWCNET)
  Extractor          <- ff_pwcnet.py:123-265   six dual-branch stages (16..196 channels) + FusionUnits
  Decoder            <- ff_pwcnet.py:267-342   cost volume -> DenseNet-style conv stack -> flow
  Refiner            <- ff_pwcnet.py:345-368   dilated context network
  backwarp           <- ff_pwcnet.py:27-46

What differs from the reference (values agree to rounding; tested against the unmodified FF_PWCNET on the GPU):
  * ``leaky_relu(FunctionCorrelation(one, two), 0.1)`` (ff_pwcnet.py:317,325) is ONE kernel (``correlation_leaky``);
  * ``backwarp`` is ONE kernel (``ffcorr_backwarp_f32``) instead of ~9 PyTorch kernels over C+1 planes; under autograd
    the reference formula is used (the kernel has no backward);
  * both images go through the extractor as one batch of 2B;
  * kernels run on torch's current stream of the tensors' device (the reference launches on CuPy's stream).
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Dict, List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import _lib
from ..correlation import FunctionCorrelation, correlation_leaky

_GRID_CACHE: Dict[tuple, tuple] = {}


def _pixel_centre_grids(h: int, w: int, device) -> tuple:
    """ff_pwcnet.py:29-31: linspace(-1 + 1/size, 1 - 1/size, size) for x and y, cached per (shape, device)."""
    key = (h, w, str(device))
    if key not in _GRID_CACHE:
        gx = torch.linspace(-1.0 + (1.0 / w), 1.0 - (1.0 / w), w).to(device)
        gy = torch.linspace(-1.0 + (1.0 / h), 1.0 - (1.0 / h), h).to(device)
        _GRID_CACHE[key] = (gx.contiguous(), gy.contiguous())
    return _GRID_CACHE[key]


def backwarp_reference_formula(ten_input: torch.Tensor, ten_flow: torch.Tensor) -> torch.Tensor:
    """The reference's op sequence (ff_pwcnet.py:27-46), differentiable; used under autograd and by the tests."""
    b, _, h, w = ten_flow.shape
    gx, gy = _pixel_centre_grids(h, w, ten_flow.device)
    grid = torch.cat([gx.view(1, 1, 1, w).expand(1, 1, h, w), gy.view(1, 1, h, 1).expand(1, 1, h, w)], 1)
    flow = torch.cat([ten_flow[:, 0:1] / ((ten_input.shape[3] - 1.0) / 2.0), ten_flow[:, 1:2] / ((ten_input.shape[2] - 1.0) / 2.0)], 1)
    x = torch.cat([ten_input, ten_flow.new_ones([b, 1, h, w])], 1)
    out = F.grid_sample(input=x, grid=(grid + flow).permute(0, 2, 3, 1), mode="bilinear", padding_mode="zeros", align_corners=False)
    mask = (out[:, -1:] > 0.999).to(out.dtype)
    return out[:, :-1] * mask


def backwarp(ten_input: torch.Tensor, ten_flow: torch.Tensor, flow_scale: float = 1.0) -> torch.Tensor:
    """``backwarp(tenInput, tenFlow * flow_scale)`` of the reference.  Inference on CUDA: one kernel."""
    if torch.is_grad_enabled() and (ten_input.requires_grad or ten_flow.requires_grad):
        return backwarp_reference_formula(ten_input, ten_flow * flow_scale)
    if not (ten_input.is_cuda and ten_flow.is_cuda):
        raise NotImplementedError("backwarp has no CPU implementation here (the reference's needs CUDA too: .cuda() at ff_pwcnet.py:33)")
    x = ten_input.float().contiguous()
    fl = ten_flow.float().contiguous()
    b, c, h, w = x.shape
    if fl.shape != (b, 2, h, w):
        raise ValueError(f"flow {tuple(fl.shape)} does not match input {tuple(x.shape)}")
    gx, gy = _pixel_centre_grids(h, w, x.device)
    out = torch.empty_like(x)
    with _lib.on_device(x, fl) as stream:
        _lib.check(_lib.lib().ffcorr_backwarp_f32(x.data_ptr(), fl.data_ptr(), gx.data_ptr(), gy.data_ptr(), out.data_ptr(),
                                                  b, c, h, w, float(flow_scale), stream), "ffcorr_backwarp_f32")
    return out


def _lrelu_stack(cin: int, cout: int) -> nn.Sequential:
    """conv3x3 stride 2, then two conv3x3 stride 1, each followed by LeakyReLU(0.1) (keys .0 .2 .4)."""
    act = lambda: nn.LeakyReLU(inplace=False, negative_slope=0.1)
    return nn.Sequential(nn.Conv2d(cin, cout, 3, 2, 1), act(), nn.Conv2d(cout, cout, 3, 1, 1), act(), nn.Conv2d(cout, cout, 3, 1, 1), act())


class Conv1x1(nn.Module):
    """parallel_fusion.py:80-88: out = q + conv1x1(v)."""

    def __init__(self, ch: int):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 1)

    def forward(self, q, v):
        return q + self.conv(v)


class FusionUnit(nn.Module):
    """parallel_fusion.py:91-146 for fusion_type '1x1conv'."""

    def __init__(self, ch: int, bidirectional: bool = True):
        super().__init__()
        self.mask2img = Conv1x1(ch)
        self.img2mask = Conv1x1(ch) if bidirectional else None

    def forward(self, mask, img):
        img_out = self.mask2img(img, mask)
        return (self.img2mask(mask, img) if self.img2mask is not None else mask), img_out


_STAGES = ("One", "Two", "Thr", "Fou", "Fiv", "Six")
_WIDTHS = (16, 32, 64, 96, 128, 196)


class Extractor(nn.Module):
    def __init__(self):
        super().__init__()
        cin = 3
        for i, (tag, ch) in enumerate(zip(_STAGES, _WIDTHS)):
            setattr(self, f"net{tag}", _lrelu_stack(cin, ch))
            setattr(self, f"mask_net{tag}", _lrelu_stack(cin, ch))
            setattr(self, f"fusion{i + 1}", FusionUnit(ch, bidirectional=(i < 5)))
            cin = ch

    def forward(self, x, mask) -> List[torch.Tensor]:
        feats = []
        for i, tag in enumerate(_STAGES):
            x = getattr(self, f"net{tag}")(x)
            mask = getattr(self, f"mask_net{tag}")(mask)
            mask, x = getattr(self, f"fusion{i + 1}")(mask, x)
            feats.append(x)
        return feats


_DEC_IN = {2: 81 + 32 + 2 + 2, 3: 81 + 64 + 2 + 2, 4: 81 + 96 + 2 + 2, 5: 81 + 128 + 2 + 2, 6: 81}
_BACKWARP_SCALE = {2: 5.0, 3: 2.5, 4: 1.25, 5: 0.625}          # ff_pwcnet.py:277 (indexed by intLevel + 1 there)


class Decoder(nn.Module):
    def __init__(self, level: int):
        super().__init__()
        cur = _DEC_IN[level]
        self.level = level
        if level < 6:
            prev = _DEC_IN[level + 1]
            self.netUpflow = nn.ConvTranspose2d(2, 2, 4, 2, 1)
            self.netUpfeat = nn.ConvTranspose2d(prev + 128 + 128 + 96 + 64 + 32, 2, 4, 2, 1)
            self.fltBackwarp = _BACKWARP_SCALE[level]
        act = lambda: nn.LeakyReLU(inplace=False, negative_slope=0.1)
        grow = 0
        for tag, ch in zip(_STAGES[:5], (128, 128, 96, 64, 32)):
            setattr(self, f"net{tag}", nn.Sequential(nn.Conv2d(cur + grow, ch, 3, 1, 1), act()))
            grow += ch
        self.netSix = nn.Sequential(nn.Conv2d(cur + grow, 2, 3, 1, 1))

    def forward(self, one, two, previous: Optional[dict]):
        if previous is None:
            feat = correlation_leaky(one, two, 0.1)                               # ff_pwcnet.py:317
        else:
            flow = self.netUpflow(previous["tenFlow"])
            upfeat = self.netUpfeat(previous["tenFeat"])
            volume = correlation_leaky(one, backwarp(two, flow, self.fltBackwarp), 0.1)   # ff_pwcnet.py:325
            feat = torch.cat([volume, one, flow, upfeat], 1)
        for tag in _STAGES[:5]:
            feat = torch.cat([getattr(self, f"net{tag}")(feat), feat], 1)
        return {"tenFlow": self.netSix(feat), "tenFeat": feat}


class Refiner(nn.Module):
    def __init__(self):
        super().__init__()
        act = lambda: nn.LeakyReLU(inplace=False, negative_slope=0.1)
        cin = 81 + 32 + 2 + 2 + 128 + 128 + 96 + 64 + 32
        layers = []
        for cout, dil in ((128, 1), (128, 2), (128, 4), (96, 8), (64, 16), (32, 1)):
            layers += [nn.Conv2d(cin, cout, 3, 1, dil, dil), act()]
            cin = cout
        layers.append(nn.Conv2d(cin, 2, 3, 1, 1, 1))
        self.netMain = nn.Sequential(*layers)

    def forward(self, x):
        return self.netMain(x)


class FocusPWC(nn.Module):
    """ff_pwcnet.py:113-434 (FF_PWCNET, 'parallel' / '1x1conv' / 'point').  Inputs in the reference's range (no
    rescaling happens in its forward)."""

    def __init__(self, mask_channel: int = 3, cfg=None):
        super().__init__()
        self.cfg = cfg or SimpleNamespace(TRAIN=SimpleNamespace(MASK_MODAL="point", MASK_CHANNEL=mask_channel))
        self.mask_channel = mask_channel
        self.netExtractor = Extractor()
        self.netTwo, self.netThr, self.netFou, self.netFiv, self.netSix = (Decoder(l) for l in (2, 3, 4, 5, 6))
        self.netRefiner = Refiner()

    @staticmethod
    def preprocess(*tensors):
        """ff_pwcnet.py:390-402: bilinear resize to multiples of 64."""
        h, w = tensors[0].shape[-2:]
        nh, nw = int(math.floor(math.ceil(h / 64.0) * 64.0)), int(math.floor(math.ceil(w / 64.0) * 64.0))
        return [F.interpolate(input=t, size=(nh, nw), mode="bilinear", align_corners=False) for t in tensors], (h, w, nh, nw)

    def forward(self, tenOne, tenTwo, mask1, mask2=None, test_mode: bool = False):
        if mask2 is None:
            mask2 = mask1
        (tenOne, tenTwo, mask1, mask2), (h, w, nh, nw) = self.preprocess(tenOne, tenTwo, mask1, mask2)
        if mask1.shape[1] != 1:                                               # init_mask 'point', ff_pwcnet.py:69-77
            raise ValueError("point masks are single-channel")
        if self.mask_channel != 1:
            mask1 = mask1.repeat(1, self.mask_channel, 1, 1)
        mask2 = torch.ones_like(mask1) * 255
        b = tenOne.shape[0]
        feats = self.netExtractor(torch.cat([tenOne, tenTwo], 0), torch.cat([mask1, mask2], 0))
        one = [f[:b].contiguous() for f in feats]
        two = [f[b:].contiguous() for f in feats]
        flows = []
        est = self.netSix(one[-1], two[-1], None)
        flows.insert(0, est["tenFlow"])
        for dec, k in ((self.netFiv, -2), (self.netFou, -3), (self.netThr, -4), (self.netTwo, -5)):
            est = dec(one[k], two[k], est)
            if dec is self.netTwo:
                est["tenFlow"] = est["tenFlow"] + self.netRefiner(est["tenFeat"])
            flows.insert(0, est["tenFlow"])
        if test_mode:
            out = F.interpolate(input=est["tenFlow"], size=(h, w), mode="bilinear", align_corners=False)
            scale = out.new_tensor([w / nw, h / nh]).view(1, 2, 1, 1)       # ff_pwcnet.py:429-430
            return out * scale
        return flows
