"""Host-model tests on CPU: the re-written FocusRAFT (PyTorch) + the ORACLE CorrBlock must
reproduce the reference FF_RAFT_FUSION outputs recorded in tests/golden/ffraft_e2e.npz.
This validates the caller of the hot path independently of the CUDA kernels."""
import os

import numpy as np
import pytest
import torch

from oracle import corr_oracle as co
from weights import fill_state_dict, synthetic_pair

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ffraft_e2e.npz")


class OracleCorrBlock:
    """torch-in / torch-out adapter of the numpy oracle (test-only)."""

    def __init__(self, fmap1, fmap2, num_levels=4, radius=4):
        self.blk = co.CorrBlock(fmap1.detach().cpu().numpy(), fmap2.detach().cpu().numpy(), num_levels, radius)
        self.device = fmap1.device

    def __call__(self, coords):
        return torch.from_numpy(self.blk(coords.detach().cpu().numpy())).to(self.device)


def make_model(device="cpu"):
    from focusflow_official_b200.host import FocusRAFT

    model = FocusRAFT()
    sd = model.state_dict()
    fill_state_dict(sd, seed=1234)
    model.load_state_dict(sd, strict=True)
    return model.to(device).eval()


def test_state_dict_keys_match_reference():
    g = np.load(GOLD)
    model = make_model()
    assert sorted(model.state_dict().keys()) == list(g["state_keys"])
    assert sum(p.numel() for p in model.parameters()) == 7662272


def test_host_model_with_oracle_corr_matches_reference_cpu():
    g = np.load(GOLD)
    model = make_model()
    model.flow_net.corr_block = OracleCorrBlock
    b, hh, ww, iters = [int(v) for v in g["b_shape"]]
    im1, im2, m1, m2 = synthetic_pair(b, hh, ww, seed=1234 + b)
    with torch.no_grad():
        lo, up = model(im1, im2, m1, m2, raft_iters=iters, test_mode=True)
    epe = torch.linalg.norm(up - torch.from_numpy(g["b_flow_up"]), dim=1)
    assert float(epe.max()) <= 1e-2, float(epe.max())
    assert float(epe.mean()) <= 1e-3
    assert np.abs(lo.numpy() - g["b_flow_lo"]).max() <= 2e-3


def test_training_mode_returns_all_predictions():
    model = make_model()
    model.flow_net.corr_block = OracleCorrBlock
    im1, im2, m1, m2 = synthetic_pair(1, 128, 128, seed=3)
    with torch.no_grad():
        preds = model(im1, im2, m1, m2, raft_iters=3, test_mode=False)
    assert len(preds) == 3 and preds[0].shape == (1, 2, 128, 128)


def test_b200_corrblock_refuses_cpu():
    model = make_model()
    im1, im2, m1, m2 = synthetic_pair(1, 128, 128, seed=3)
    with pytest.raises(NotImplementedError):
        with torch.no_grad():
            model(im1, im2, m1, m2, raft_iters=1, test_mode=True)


def test_inference_plumbing_helpers_are_the_same_functions():
    """The inference-only shortcuts of host/focusraft.py, checked on CPU through their helpers: eval-mode BatchNorm folded
    into the convolution in front of it, two same-shape convolutions stacked along the output channels, lerp as the GRU
    state update; and the caches follow in-place weight updates."""
    import torch.nn.functional as F

    from focusflow_official_b200.host import focusraft as FR

    torch.manual_seed(0)
    conv, bn = torch.nn.Conv2d(5, 7, 3, padding=1), torch.nn.BatchNorm2d(7).eval()
    with torch.no_grad():
        bn.running_mean.normal_(); bn.running_var.uniform_(0.5, 2.0); bn.weight.normal_(); bn.bias.normal_()
        x = torch.randn(2, 5, 9, 11)
        w, b = FR._folded_bn(conv, bn)
        assert torch.allclose(F.conv2d(x, w, b, padding=1), bn(conv(x)), atol=2e-6)
        assert FR._folded_bn(conv, bn)[0] is w                       # cached
        bn.running_mean.add_(1.0)                                    # in-place update -> re-folded
        w2, b2 = FR._folded_bn(conv, bn)
        assert w2 is not w and torch.allclose(F.conv2d(x, w2, b2, padding=1), bn(conv(x)), atol=2e-6)

        ca, cb = torch.nn.Conv2d(6, 4, (1, 5), padding=(0, 2)), torch.nn.Conv2d(6, 4, (1, 5), padding=(0, 2))
        y = torch.randn(2, 6, 8, 10)
        ws, bs = FR._stacked(ca, cb)
        both = F.conv2d(y, ws, bs, padding=(0, 2))
        assert torch.allclose(both[:, :4], ca(y), atol=1e-6) and torch.allclose(both[:, 4:], cb(y), atol=1e-6)
        ca.weight.mul_(2.0)
        assert torch.allclose(F.conv2d(y, *FR._stacked(ca, cb), padding=(0, 2))[:, :4], ca(y), atol=1e-6)

        h, q, z = torch.randn(3, 4), torch.randn(3, 4), torch.rand(3, 4)
        assert torch.allclose(torch.lerp(h, q, z), (1 - z) * h + z * q, atol=1e-6)

        enc = FR.MotionEncoder(4, 4)
        flow, corr = torch.randn(1, 2, 6, 8), torch.randn(1, 324, 6, 8)
        assert torch.equal(enc(flow, corr), torch.cat([enc.features(flow, corr), flow], dim=1))
