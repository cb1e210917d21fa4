"""FocusRAFT host model (PyTorch) around the B200 correlation kernels.

Re-written caller of the hot path; mirrors, module name by module name, the configuration every
reference experiment uses (`FUSION: parallel`, `FUSION_TYPE: 1x1conv`, `FUSE_CNET: true`,
`SMALL: false`, `MASK_MODAL: point`):

  FocusRAFT            <- FF_RAFT_Core/ff_raft.py:75-160   (FF_RAFT_FUSION)
  RAFTBody             <- FF_RAFT_Core/raft.py:41-236      (RAFT)
  CCEEncoder           <- FF_RAFT_Core/parallel_fusion.py:153-247 + extractor.py:118-192
  UpdateBlock & co.    <- FF_RAFT_Core/update.py:6-135

Only `CorrBlock` differs: it is `focusflow_official_b200.CorrBlock` (sm_100a kernels) instead of
`FF_RAFT_Core/corr.py`.  Host-side restructurings that do not change the math:
  * both images go through the feature encoder as ONE batch of 2B (instance norm is per sample);
  * in test mode the convex upsampling runs once, on the last iteration (the reference computes
    it every iteration and returns only the last, raft.py:226-234);
  * `corr_block=` lets tests/benchmarks swap in the CPU oracle for the same host code.
"""
from __future__ import annotations

import weakref
from types import SimpleNamespace
from typing import Callable, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..corr import AlternateCorrBlock as B200AlternateCorrBlock
from ..corr import CorrBlock as B200CorrBlock
from ..corr import coords_grid


def _conv_relu(conv: nn.Conv2d, x: torch.Tensor) -> torch.Tensor:
    """relu(conv(x) + bias).  Inference on CUDA: ONE cuDNN ConvBiasAct call instead of conv + bias-add + relu
    kernels (the broadcast bias add alone was 18 % of a FocusRAFT step).  Same math as the reference's
    ``relu(conv(x))``; under autograd the plain composition is used."""
    if x.is_cuda and not torch.is_grad_enabled() and conv.bias is not None and conv.padding_mode == "zeros":
        return torch.cudnn_convolution_relu(x, conv.weight, conv.bias, conv.stride, conv.padding, conv.dilation, conv.groups)
    return F.relu(conv(x))


_folded_cache: "weakref.WeakKeyDictionary" = weakref.WeakKeyDictionary()


def _folded_bn(conv: nn.Conv2d, norm: nn.BatchNorm2d):
    """Eval-mode BatchNorm folded into the convolution in front of it: w' = w * g / sqrt(var + eps),
    b' = (b - mean) * g / sqrt(var + eps) + beta.  Cached until any of the six tensors changes."""
    ts = (conv.weight, conv.bias, norm.weight, norm.bias, norm.running_mean, norm.running_var)
    key = tuple((t.data_ptr(), t._version) for t in ts) + (conv.weight.is_contiguous(memory_format=torch.channels_last),)
    hit = _folded_cache.get(conv)
    if hit is not None and hit[0] == key:
        return hit[1], hit[2]
    scale = norm.weight * torch.rsqrt(norm.running_var + norm.eps)
    w = conv.weight * scale.view(-1, 1, 1, 1)
    if key[-1]:
        w = w.contiguous(memory_format=torch.channels_last)
    b = (conv.bias - norm.running_mean) * scale + norm.bias
    _folded_cache[conv] = (key, w, b)
    return w, b


_stacked_cache: "weakref.WeakKeyDictionary" = weakref.WeakKeyDictionary()


def _stacked(a: nn.Conv2d, b: nn.Conv2d):
    """Weights and biases of two convolutions of the same geometry stacked along the output channels (cached until one of
    the four tensors changes): conv(x, cat(wa, wb)) == cat(conv_a(x), conv_b(x))."""
    ts = (a.weight, a.bias, b.weight, b.bias)
    key = tuple((t.data_ptr(), t._version) for t in ts) + (a.weight.is_contiguous(memory_format=torch.channels_last),)
    hit = _stacked_cache.get(a)
    if hit is not None and hit[0] == key:
        return hit[1], hit[2]
    w = torch.cat([a.weight, b.weight], dim=0)
    if key[-1]:
        w = w.contiguous(memory_format=torch.channels_last)
    bias = torch.cat([a.bias, b.bias], dim=0)
    _stacked_cache[a] = (key, w, bias)
    return w, bias


def _conv_norm(conv: nn.Conv2d, norm: nn.Module, x: torch.Tensor, relu: bool = False) -> torch.Tensor:
    """[relu(] norm(conv(x)) [)] with fewer kernels at inference on CUDA:
    InstanceNorm (no affine, per-sample statistics) subtracts the channel mean, so the conv bias cancels exactly and its
    broadcast add is dropped; eval-mode BatchNorm is an affine map per channel, so it is FOLDED into the convolution's
    weights and bias (one cuDNN ConvBiasAct call instead of conv + bias add + batch-norm + relu).  Same function as the
    reference's composition up to fp32 rounding of the folded weights; under autograd the plain composition is used."""
    if x.is_cuda and not torch.is_grad_enabled() and conv.bias is not None and conv.padding_mode == "zeros":
        if isinstance(norm, nn.InstanceNorm2d) and not norm.affine and not norm.track_running_stats:
            y = norm(F.conv2d(x, conv.weight, None, conv.stride, conv.padding, conv.dilation, conv.groups))
            return F.relu(y, inplace=True) if relu else y
        if isinstance(norm, nn.BatchNorm2d) and not norm.training and norm.track_running_stats and norm.affine:
            w, b = _folded_bn(conv, norm)
            if relu:
                return torch.cudnn_convolution_relu(x, w, b, conv.stride, conv.padding, conv.dilation, conv.groups)
            return F.conv2d(x, w, b, conv.stride, conv.padding, conv.dilation, conv.groups)
    y = norm(conv(x))
    return F.relu(y, inplace=True) if relu else y


def _norm(kind: str, ch: int) -> nn.Module:
    if kind == "instance":
        return nn.InstanceNorm2d(ch)
    if kind == "batch":
        return nn.BatchNorm2d(ch)
    if kind == "group":
        return nn.GroupNorm(num_groups=ch // 8, num_channels=ch)
    if kind == "none":
        return nn.Sequential()
    raise ValueError(f"unknown norm {kind!r}")


class ResidualBlock(nn.Module):
    """extractor.py:6-56 (keys: conv1, conv2, norm1, norm2[, norm3, downsample.{0,1}])."""

    def __init__(self, cin: int, cout: int, norm: str, stride: int = 1):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1, stride=stride)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.relu = nn.ReLU(inplace=True)
        self.norm1 = _norm(norm, cout)
        self.norm2 = _norm(norm, cout)
        self.downsample = None
        if stride != 1:
            self.norm3 = _norm(norm, cout)  # registered twice on purpose: reference checkpoints carry both keys
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride=stride), self.norm3)

    def forward(self, x):
        y = _conv_norm(self.conv1, self.norm1, x, relu=True)
        y = _conv_norm(self.conv2, self.norm2, y, relu=True)
        if self.downsample is not None:
            x = _conv_norm(self.downsample[0], self.downsample[1], x)
        return self.relu(x + y)


def _stage(cin: int, cout: int, norm: str, stride: int) -> nn.Sequential:
    return nn.Sequential(ResidualBlock(cin, cout, norm, stride), ResidualBlock(cout, cout, norm, 1))


class Conv1x1(nn.Module):
    """parallel_fusion.py:87-95: out = q + conv1x1(v)."""

    def __init__(self, ch: int):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 1)

    def forward(self, q, v):
        return q + self.conv(v)


class FusionUnit(nn.Module):
    """parallel_fusion.py:98-150 for fusion_type '1x1conv'."""

    def __init__(self, ch: int, bidirectional: bool = True):
        super().__init__()
        self.mask2img = Conv1x1(ch)
        self.img2mask = Conv1x1(ch) if bidirectional else None

    def forward(self, mask, img):
        img_out = self.mask2img(img, mask)  # both directions read the un-updated inputs
        mask_out = self.img2mask(mask, img) if self.img2mask is not None else mask
        return mask_out, img_out


class CCEEncoder(nn.Module):
    """FFE (image branch) + CFE (mask branch) + 5 fusion units; 1/8 resolution output."""

    def __init__(self, img_channel: int = 3, mask_channel: int = 3, output_dim: int = 256, norm_fn: str = "instance"):
        super().__init__()
        self.norm_fn = norm_fn
        # image branch (same attribute names as the reference BasicEncoder)
        self.norm1 = _norm(norm_fn, 64)
        self.conv1 = nn.Conv2d(img_channel, 64, 7, stride=2, padding=3)
        self.relu1 = nn.ReLU(inplace=True)
        self.layer1 = _stage(64, 64, norm_fn, 1)
        self.layer2 = _stage(64, 96, norm_fn, 2)
        self.layer3 = _stage(96, 128, norm_fn, 2)
        self.conv2 = nn.Conv2d(128, output_dim, 1)
        # mask branch
        self.mask_norm1 = _norm(norm_fn, 64)
        self.mask_conv1 = nn.Conv2d(mask_channel, 64, 7, stride=2, padding=3)
        self.mask_relu1 = nn.ReLU(inplace=True)
        self.mask_layer1 = _stage(64, 64, norm_fn, 1)
        self.mask_layer2 = _stage(64, 96, norm_fn, 2)
        self.mask_layer3 = _stage(96, 128, norm_fn, 2)
        self.mask_conv2 = nn.Conv2d(128, output_dim, 1)
        self.fusion1 = FusionUnit(64)
        self.fusion2 = FusionUnit(64)
        self.fusion3 = FusionUnit(96)
        self.fusion4 = FusionUnit(128)
        self.fusion5 = FusionUnit(output_dim, bidirectional=False)
        for m in self.modules():  # parallel_fusion.py:193-200
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, (nn.BatchNorm2d, nn.GroupNorm)):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def forward(self, x, mask):
        mask = _conv_norm(self.mask_conv1, self.mask_norm1, mask, relu=True)
        x = _conv_norm(self.conv1, self.norm1, x, relu=True)
        mask, x = self.fusion1(mask, x)
        mask, x = self.fusion2(self.mask_layer1(mask), self.layer1(x))
        mask, x = self.fusion3(self.mask_layer2(mask), self.layer2(x))
        mask, x = self.fusion4(self.mask_layer3(mask), self.layer3(x))
        mask, x = self.fusion5(self.mask_conv2(mask), self.conv2(x))
        return x


class FlowHead(nn.Module):
    def __init__(self, cin: int = 128, hidden: int = 256):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, hidden, 3, padding=1)
        self.conv2 = nn.Conv2d(hidden, 2, 3, padding=1)
        self.relu = nn.ReLU(inplace=True)

    def forward(self, x):
        return self.conv2(_conv_relu(self.conv1, x))


class SepConvGRU(nn.Module):
    """update.py:33-60: a 1x5 GRU step followed by a 5x1 GRU step."""

    def __init__(self, hidden: int = 128, inp: int = 320):
        super().__init__()
        for tag, k, p in (("1", (1, 5), (0, 2)), ("2", (5, 1), (2, 0))):
            for gate in "zrq":
                setattr(self, f"conv{gate}{tag}", nn.Conv2d(hidden + inp, hidden, k, padding=p))

    def _step(self, h, x, cz, cr, cq):
        hx = torch.cat([h, x], dim=1)
        if hx.is_cuda and not torch.is_grad_enabled():
            # inference plumbing: the z and r gates read the same input through convolutions of the same shape -> ONE
            # convolution with the two weight sets stacked along the output channels (every output channel is computed
            # exactly as before), one sigmoid; and the state update as one lerp kernel: h + z (q - h) == (1 - z) h + z q
            w, b = _stacked(cz, cr)
            zr = torch.sigmoid(F.conv2d(hx, w, b, cz.stride, cz.padding, cz.dilation, cz.groups))
            z, r = zr[:, :cz.out_channels], zr[:, cz.out_channels:]
            q = torch.tanh(cq(torch.cat([r * h, x], dim=1)))
            return torch.lerp(h, q, z)
        z = torch.sigmoid(cz(hx))
        r = torch.sigmoid(cr(hx))
        q = torch.tanh(cq(torch.cat([r * h, x], dim=1)))
        return (1 - z) * h + z * q

    def forward(self, h, x):
        h = self._step(h, x, self.convz1, self.convr1, self.convq1)
        return self._step(h, x, self.convz2, self.convr2, self.convq2)


class MotionEncoder(nn.Module):
    """update.py:79-97; convc1 consumes the num_levels*(2r+1)^2 correlation planes."""

    def __init__(self, corr_levels: int, corr_radius: int):
        super().__init__()
        planes = corr_levels * (2 * corr_radius + 1) ** 2
        self.convc1 = nn.Conv2d(planes, 256, 1)
        self.convc2 = nn.Conv2d(256, 192, 3, padding=1)
        self.convf1 = nn.Conv2d(2, 128, 7, padding=3)
        self.convf2 = nn.Conv2d(128, 64, 3, padding=1)
        self.conv = nn.Conv2d(64 + 192, 128 - 2, 3, padding=1)

    def features(self, flow, corr, cor1=None):
        """The 126 learned motion channels (update.py:90-96 without the final concatenation with `flow`)."""
        # cor1: relu(convc1(corr)) already computed by the fused lookup (CorrBlock.lookup_conv)
        cor = _conv_relu(self.convc2, _conv_relu(self.convc1, corr) if cor1 is None else cor1)
        flo = _conv_relu(self.convf2, _conv_relu(self.convf1, flow))
        return _conv_relu(self.conv, torch.cat([cor, flo], dim=1))

    def forward(self, flow, corr, cor1=None):
        return torch.cat([self.features(flow, corr, cor1), flow], dim=1)


class UpdateBlock(nn.Module):
    """update.py:114-135 (BasicUpdateBlock)."""

    def __init__(self, corr_levels: int, corr_radius: int, hidden: int = 128):
        super().__init__()
        self.encoder = MotionEncoder(corr_levels, corr_radius)
        self.gru = SepConvGRU(hidden, 128 + hidden)
        self.flow_head = FlowHead(hidden, 256)
        self.mask = nn.Sequential(nn.Conv2d(128, 256, 3, padding=1), nn.ReLU(inplace=True), nn.Conv2d(256, 64 * 9, 1))

    def forward(self, net, inp, corr, flow, with_mask: bool = True, cor1=None):
        # cat([inp, cat([out, flow])]) (update.py:97,128-129) as one three-way concatenation: same channel order
        net = self.gru(net, torch.cat([inp, self.encoder.features(flow, corr, cor1), flow], dim=1))
        delta = self.flow_head(net)
        mask = 0.25 * self.mask[2](_conv_relu(self.mask[0], net)) if with_mask else None  # 0.25: update.py:133-134
        return net, mask, delta


def convex_upsample(flow, mask):
    """raft.py:159-170: [B,2,h,w] -> [B,2,8h,8w] as a softmax-weighted 3x3 combination."""
    b, _, h, w = flow.shape
    mask = torch.softmax(mask.view(b, 1, 9, 8, 8, h, w), dim=2)
    up = F.unfold(8 * flow, [3, 3], padding=1).view(b, 2, 9, 1, 1, h, w)
    up = torch.sum(mask * up, dim=2).permute(0, 1, 4, 2, 5, 3)
    return up.reshape(b, 2, 8 * h, 8 * w)


class RAFTBody(nn.Module):
    """raft.py:41-236 for inside_fusion='parallel', fuse_cnet=True, small=False."""

    hidden_dim = 128
    context_dim = 128
    corr_levels = 4
    corr_radius = 4

    def __init__(self, mask_channel: int = 3):
        super().__init__()
        self.fnet = CCEEncoder(3, mask_channel, 256, "instance")
        self.cnet = CCEEncoder(3, mask_channel, self.hidden_dim + self.context_dim, "batch")
        self.update_block = UpdateBlock(self.corr_levels, self.corr_radius, self.hidden_dim)
        self.corr_block: Callable = B200CorrBlock
        self.corr_precision: Optional[str] = None
        self.corr_sampler: Optional[str] = None      # None = the package default ("cuda": the reference's GPU run)
        self.corr_storage: Optional[str] = None      # None / "fp32" (default) or "fp16": opt-in half-precision pyramid
        # opt-in: the lookup and the motion encoder's first layer (convc1 + ReLU, update.py:90) in one kernel
        # (CorrBlock.lookup_conv); needs update_channels_last, inference only
        self.fuse_convc1 = False
        # raft.py:63,195-196 `alternate_corr` (config key ALT_CORR): the memory-bounded block that recomputes the
        # pyramid per lookup; inference only, like the reference's alt_cuda_corr
        self.alternate_corr = False
        # host plumbing knob: run the GRU update block in channels_last (NHWC) so cuDNN's tensor-core
        # kernels need no per-conv NCHW<->NHWC transposes; numerically the same convolutions
        self.update_channels_last = False

    def freeze_bn(self):
        for m in self.modules():
            if isinstance(m, nn.BatchNorm2d):
                m.eval()

    def _make_corr(self, fmap1, fmap2):
        # the lookup writes NHWC directly when the update block runs channels_last (its consumer convc1, update.py:90)
        kw = dict(num_levels=self.corr_levels, radius=self.corr_radius, precision=self.corr_precision,
                  sampler=self.corr_sampler, channels_last=self.update_channels_last)
        if self.alternate_corr and self.corr_block is B200CorrBlock and not torch.is_grad_enabled():
            return B200AlternateCorrBlock(fmap1, fmap2, **kw)
        if self.corr_block is B200CorrBlock:
            return B200CorrBlock(fmap1, fmap2, storage=self.corr_storage, **kw)
        return self.corr_block(fmap1, fmap2, num_levels=self.corr_levels, radius=self.corr_radius)

    def forward(self, image1, image2, mask1, mask2, iters: int = 12, flow_init=None, test_mode: bool = False):
        b = image1.shape[0]
        fmaps = self.fnet(torch.cat([image1, image2], dim=0), torch.cat([mask1, mask2], dim=0)).float()
        fmap1, fmap2 = fmaps[:b].contiguous(), fmaps[b:].contiguous()
        corr_fn = self._make_corr(fmap1, fmap2)  # raft.py:198

        cnet = self.cnet(image1, mask1)
        net, inp = torch.split(cnet, [self.hidden_dim, self.context_dim], dim=1)
        net, inp = torch.tanh(net), torch.relu(inp)

        h, w = image1.shape[2] // 8, image1.shape[3] // 8
        coords0 = coords_grid(b, h, w, image1.device)
        coords1 = coords0.clone()
        if flow_init is not None:
            coords1 = coords1 + flow_init

        cl = self.update_channels_last
        if cl:
            net = net.contiguous(memory_format=torch.channels_last)
            inp = inp.contiguous(memory_format=torch.channels_last)
        predictions = []
        flow_up = None
        convc1 = self.update_block.encoder.convc1
        fused = bool(self.fuse_convc1 and cl and hasattr(corr_fn, "supports_lookup_conv") and corr_fn.supports_lookup_conv(convc1))
        for it in range(iters):
            coords1 = coords1.detach()  # raft.py:216 -> the lookup never needs d/d coords
            corr = cor1 = None
            if fused:
                cor1 = corr_fn.lookup_conv(coords1, convc1)
            else:
                corr = corr_fn(coords1)
            flow = coords1 - coords0
            if cl:
                if corr is not None:
                    corr = corr.contiguous(memory_format=torch.channels_last)   # a no-op for the B200 block (already NHWC)
                flow = flow.contiguous(memory_format=torch.channels_last)
            need_up = (not test_mode) or it == iters - 1
            net, up_mask, delta = self.update_block(net, inp, corr, flow, with_mask=need_up, cor1=cor1)
            coords1 = coords1 + delta
            if need_up:
                flow_up = convex_upsample(coords1 - coords0, up_mask)
                predictions.append(flow_up)
        if test_mode:
            return coords1 - coords0, flow_up
        return predictions


def init_point_mask(mask1, mask_channel: int = 3):
    """ff_raft.py:31-38 ('point' modality): mask1 repeated to 3 channels, mask2 = all 255."""
    if mask1.shape[1] != 1:
        raise ValueError("point masks are single-channel")
    if mask_channel != 1:
        mask1 = mask1.repeat(1, mask_channel, 1, 1)
    return mask1, torch.full_like(mask1, 255.0)


class FocusRAFT(nn.Module):
    """ff_raft.py:75-160 (FF_RAFT_FUSION, use_fusion='parallel').  Inputs in [0, 255]."""

    def __init__(self, mask_channel: int = 3, cfg=None):
        super().__init__()
        self.cfg = cfg or SimpleNamespace(TRAIN=SimpleNamespace(MASK_MODAL="point", MASK_CHANNEL=mask_channel))
        self.mask_channel = mask_channel
        self.flow_net = RAFTBody(mask_channel)

    def forward(self, image1, image2, mask1, mask2=None, raft_iters: int = 12, flow_init=None, test_mode: bool = False):
        mask1, mask2 = init_point_mask(mask1, self.mask_channel)
        scale = lambda t: 2 * (t.contiguous() / 255.0) - 1.0  # ff_raft.py:142-145, masks included
        return self.flow_net(scale(image1), scale(image2), scale(mask1), scale(mask2), iters=raft_iters,
                             flow_init=flow_init, test_mode=test_mode)


def build_focusraft(device="cuda", seed: Optional[int] = 1234, channels_last: bool = False) -> FocusRAFT:
    if seed is not None:
        torch.manual_seed(seed)
    model = FocusRAFT().to(device)
    if channels_last:
        model = model.to(memory_format=torch.channels_last)
    return model.eval()
