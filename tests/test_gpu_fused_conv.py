"""GPU tests of the lookup fused with its consumer: ``CorrBlock.lookup_conv`` =
``relu(convc1(lookup(coords)))`` (``update.py:82-83,90``) in one launch, through ``ffcorr_lookup_convc1_tiled_f32``.

The kernel's stated arithmetic: the lookup's fp32 samples rounded to fp16 (RN, saturating), convc1's weights rounded to
fp16, products accumulated in fp32 on the tensor cores, + bias, ReLU.  fp16 keeps the 11 significant bits of the TF32
convolution the reference runs under ALLOW_TF32 (common.py:25-27), so the bar against the exact fp32 convolution is
the TF32 convolution's own error."""
import numpy as np
import pytest
import torch

from oracle import corr_oracle as co

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def ff():
    import focusflow_official_b200 as m

    return m


def _case(b, h, w, seed=3, scale=2.0, sampler="cuda", same=False):
    torch.manual_seed(seed)
    f1 = torch.randn(b, 256, h, w, device=DEV) * scale
    f2 = f1.clone() if same else torch.randn(b, 256, h, w, device=DEV) * scale
    conv = torch.nn.Conv2d(324, 256, 1).to(DEV)
    with torch.no_grad():
        conv.weight.mul_(2.0)
        conv.bias.uniform_(-0.5, 0.5)
    grid = ff().coords_grid(b, h, w, DEV)
    flow = torch.nn.functional.interpolate(torch.randn(b, 2, h // 8 + 1, w // 8 + 1, device=DEV) * 5, size=(h, w), mode="bilinear")
    coords = grid + flow
    coords[:, :, 0, 0] = -50.0                       # a window entirely outside the map
    coords[:, 0, -1, -1] = float("inf")              # "wild" coordinates: all taps zero
    coords[:, :, h // 2, w // 2] += 0.5
    blk = ff().CorrBlock(f1, f2, channels_last=True, sampler=sampler)
    return blk, conv, coords


def _expected_fp16_operands(samples, conv):
    """relu(W16 . v16 + b) in float64 from the fp32 samples [B, 324, h, w]."""
    v16 = samples.clamp(-65504.0, 65504.0).half().double()
    w16 = conv.weight.detach().reshape(256, 324).half().double()
    out = torch.einsum("oc,bchw->bohw", w16, v16) + conv.bias.detach().double()[None, :, None, None]
    return torch.relu(out)


@pytest.mark.parametrize("shape", [(1, 46, 62), (2, 24, 40), (1, 33, 45), (3, 8, 8), (1, 47, 156)])
@pytest.mark.parametrize("sampler", ["cuda", "cpu"])
def test_lookup_conv_is_the_stated_arithmetic(shape, sampler):
    b, h, w = shape
    blk, conv, coords = _case(b, h, w, sampler=sampler)
    with torch.no_grad():
        samples = blk(coords)
        got = blk.lookup_conv(coords, conv)
        assert got.shape == (b, 256, h, w) and got.is_contiguous(memory_format=torch.channels_last) or got.shape[2] * got.shape[3] == 1
        assert torch.isfinite(got).all()
        exp = _expected_fp16_operands(samples, conv)
        scale = exp.abs().max().item()
        # same operands, fp32 accumulation of 324 products: a few ulp of the largest partial sum
        assert (got.double() - exp).abs().max().item() <= 2e-5 * scale, (shape, (got.double() - exp).abs().max().item(), scale)
        # against the exact fp32 convolution: no worse than the reference's TF32 convolution
        before = torch.backends.cudnn.allow_tf32
        try:
            torch.backends.cudnn.allow_tf32 = False
            exact = torch.relu(torch.nn.functional.conv2d(samples, conv.weight, conv.bias))
            torch.backends.cudnn.allow_tf32 = True
            tf32 = torch.relu(torch.nn.functional.conv2d(samples, conv.weight, conv.bias))
        finally:
            torch.backends.cudnn.allow_tf32 = before
        err = (got - exact).abs().max().item()
        err_tf32 = (tf32 - exact).abs().max().item()
        assert err <= 1e-3 * scale, (shape, err, scale)
        assert err <= 1.5 * err_tf32 + 1e-5 * scale, (shape, err, err_tf32)


def test_lookup_conv_against_the_cpu_oracle_lookup():
    """Independent of this repo's lookup kernel: oracle lookup (corr.py:29-50 restated) + numpy convolution."""
    b, h, w = 1, 24, 40
    blk, conv, coords = _case(b, h, w, sampler="cpu")
    with torch.no_grad():
        pyr = [lv[:, 0].cpu().numpy() for lv in blk.corr_pyramid]
        ref = co.lookup(pyr, coords.cpu().numpy(), 4)                     # [B, 324, h, w]
        wt = conv.weight.detach().reshape(256, 324).cpu().numpy().astype(np.float64)
        exp = np.maximum(np.einsum("oc,bchw->bohw", wt, ref.astype(np.float64)) + conv.bias.detach().cpu().numpy()[None, :, None, None], 0.0)
        got = blk.lookup_conv(coords, conv).cpu().numpy()
    assert np.abs(got - exp).max() <= 1e-3 * np.abs(exp).max()


def test_lookup_conv_saturates_instead_of_overflowing():
    blk, conv, _ = _case(1, 16, 24, scale=80.0, same=True)                # self-correlations of ~1e5: beyond fp16
    coords = ff().coords_grid(1, 16, 24, DEV)
    with torch.no_grad():
        samples = blk(coords)
        assert samples.abs().max().item() > 65504.0
        got = blk.lookup_conv(coords, conv)
        assert torch.isfinite(got).all()
        exp = _expected_fp16_operands(samples, conv)
        assert (got.double() - exp).abs().max().item() <= 2e-5 * exp.abs().max().item()


def test_lookup_conv_repacks_when_the_weights_change_and_refuses_what_it_cannot_do():
    blk, conv, coords = _case(1, 16, 24)
    with torch.no_grad():
        a = blk.lookup_conv(coords, conv)
        conv.weight.mul_(0.5)                                              # in-place update bumps the version counter
        conv.bias.zero_()
        b2 = blk.lookup_conv(coords, conv)
        exp = _expected_fp16_operands(blk(coords), conv)
        assert (b2.double() - exp).abs().max().item() <= 2e-5 * exp.abs().max().item()
        assert not torch.equal(a, b2)
        assert blk.supports_lookup_conv(conv)
        assert not blk.supports_lookup_conv(torch.nn.Conv2d(324, 128, 1).to(DEV))
        assert not blk.supports_lookup_conv(torch.nn.Conv2d(324, 256, 1, bias=False).to(DEV))
        small = ff().CorrBlock(torch.randn(1, 32, 16, 24, device=DEV), torch.randn(1, 32, 16, 24, device=DEV), num_levels=3, radius=3)
        assert not small.supports_lookup_conv(conv)
        with pytest.raises(ValueError):
            small.lookup_conv(coords, conv)
        with pytest.raises(ValueError):
            blk.lookup_conv(coords[:, :, :8], conv)
    with torch.enable_grad():
        assert not blk.supports_lookup_conv(conv)                          # inference only


def test_host_model_with_fused_convc1_gives_the_same_flow():
    """FocusRAFT with the fused first layer of the motion encoder: same flow as the unfused host model up to the rounding
    differences between two TF32-grade convolutions (accumulation order)."""
    from test_host_model import make_model
    from weights import synthetic_pair

    model = make_model().to(DEV).eval()
    model.flow_net.update_block.to(memory_format=torch.channels_last)
    model.flow_net.update_channels_last = True
    im1, im2, m1, _ = (x.to(DEV) for x in synthetic_pair(2, 128, 192, seed=3))
    before = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        torch.backends.cudnn.allow_tf32 = True                             # the reference's run setting (common.py:25-27)
        with torch.no_grad():
            ref_lo, ref = model(im1, im2, m1, None, raft_iters=12, test_mode=True)
            model.flow_net.fuse_convc1 = True
            got_lo, got = model(im1, im2, m1, None, raft_iters=12, test_mode=True)
            torch.backends.cudnn.allow_tf32 = False
            model.flow_net.fuse_convc1 = False
            exact_lo, exact = model(im1, im2, m1, None, raft_iters=12, test_mode=True)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = before
    epe = (got - ref).norm(dim=1)
    drift = (ref - exact).norm(dim=1)                                      # what TF32 convolutions alone do to the flow
    print(f"fused vs unfused EPE mean {epe.mean().item():.2e} max {epe.max().item():.2e}; "
          f"TF32-vs-fp32 host convs: mean {drift.mean().item():.2e} max {drift.max().item():.2e}")
    assert torch.isfinite(got).all()
    assert epe.max().item() <= max(2.0 * drift.max().item(), 2e-3)


def test_packed_weight_layout():
    """ffcorr_pack_convc1_weight: W[co, level*81 + a*9 + bb] (corr.py:37-43 channel order) lands at W'[co, level*90 + bb*10 + a]
    as fp16 (RN); every other slot of the 384 is zero."""
    from focusflow_official_b200.corr import _packed_convc1

    conv = torch.nn.Conv2d(324, 256, 1).to(DEV)
    packed = _packed_convc1(conv).view(torch.float16).view(256, 384).cpu().numpy()
    w = conv.weight.detach().reshape(256, 324).cpu().numpy()
    exp = np.zeros((256, 384), np.float16)
    for level in range(4):
        for a in range(9):
            for bb in range(9):
                exp[:, level * 90 + bb * 10 + a] = w[:, level * 81 + a * 9 + bb].astype(np.float16)
    assert np.array_equal(packed, exp)
    assert _packed_convc1(conv) is _packed_convc1(conv)          # cached until the weights change
