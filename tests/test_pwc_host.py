"""FF-PWC caller of the local cost volume (BASELINE config 3): host model, backwarp kernel.

The reference's FF_PWCNET cannot run on CPU at all (backwarp calls .cuda(), ff_pwcnet.py:33; the cost volume raises,
correlation.py:320-321), so value parity is a GPU test: the UNMODIFIED reference model (from /root/reference or
baseline/_ref) with a stand-in `correlation` module -- the closed-form torch restatement of correlation.py:46-98, itself
pinned to the reference's CUDA kernels in test_oracle_golden.py -- against this repo's host model with the B200 kernels,
same deterministic weights."""
import os
import sys

import numpy as np
import pytest
import torch

from weights import PWC_GAINS, fill_state_dict, synthetic_pair

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def torch_correlation(tenOne, tenTwo):
    """correlation.py:46-98 in closed form: mean over channels of one * shifted(two), 81 displacements."""
    b, c, h, w = tenOne.shape
    pad = torch.nn.functional.pad(tenTwo, (4, 4, 4, 4))
    out = [(tenOne * pad[:, :, 4 + dy:4 + dy + h, 4 + dx:4 + dx + w]).mean(1) for dy in range(-4, 5) for dx in range(-4, 5)]
    return torch.stack(out, 1)


def host_pwc(device="cpu"):
    from focusflow_official_b200.host import FocusPWC

    model = FocusPWC()
    sd = model.state_dict()
    fill_state_dict(sd, seed=99, gains=PWC_GAINS)
    model.load_state_dict(sd, strict=True)
    return model.to(device).eval()


def test_state_dict_is_the_references():
    from oracle import reference_loader as RL

    model = host_pwc()
    assert len(model.state_dict()) == 184 and sum(p.numel() for p in model.parameters()) == 11141314
    if RL.reference_root("ff-pwcnet") is None:
        pytest.skip("reference sources not installed")
    ref, _ = RL.load_ff_pwc(torch_correlation)
    rsd, sd = ref.state_dict(), model.state_dict()
    assert sorted(rsd) == sorted(sd) and all(rsd[k].shape == sd[k].shape for k in rsd)


def test_cpu_is_refused():
    from focusflow_official_b200.host import backwarp

    with pytest.raises(NotImplementedError):
        backwarp(torch.zeros(1, 4, 8, 8), torch.zeros(1, 2, 8, 8))


@pytest.mark.gpu
@pytest.mark.parametrize("shape,scale", [((2, 32, 28, 64), 5.0), ((1, 96, 14, 32), 1.25), ((2, 20, 11, 14), 0.625), ((1, 128, 7, 16), 2.5),
                                         ((4, 32, 112, 256), 2.5), ((1, 196, 9, 20), 1.0), ((3, 7, 5, 70), 1.0)])
def test_backwarp_kernel_vs_the_reference_formula(shape, scale):
    """ffcorr_backwarp_f32 == ff_pwcnet.py:27-46 run through torch's CUDA kernels: flows that stay inside, leave the
    image (validity mask), land exactly on pixel centres, and are non-finite."""
    from focusflow_official_b200.host import backwarp
    from focusflow_official_b200.host.focuspwc import backwarp_reference_formula

    b, c, h, w = shape
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn(shape, device="cuda", generator=g)
    for sigma in (0.0, 0.3, 2.0, 20.0):
        flow = torch.randn(b, 2, h, w, device="cuda", generator=g) * sigma
        if sigma == 20.0:
            flow[0, 0, 0, 0] = float("inf")
            flow[0, 1, 1, 1] = float("nan")
        ref = backwarp_reference_formula(x, flow * scale)
        got = backwarp(x, flow, scale)
        ok = torch.isfinite(ref)
        assert torch.isfinite(got).all()
        err = float((got - ref)[ok].abs().max())
        assert err <= 1e-5 * max(1.0, float(ref[ok].abs().max())), (shape, sigma, err)
        # the validity mask must agree pixel for pixel (a flipped 0.999 threshold would show as an O(1) error)
        assert int(((got == 0).all(1) != (ref == 0).all(1))[ok.all(1)].sum()) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("hw,batch", [((128, 192), 2), ((436, 1024), 1)])
def test_host_model_matches_the_unmodified_reference_on_the_gpu(hw, batch):
    from oracle import reference_loader as RL

    if RL.reference_root("ff-pwcnet") is None:
        pytest.skip("reference sources not installed under baseline/_ref")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    model = host_pwc("cuda")
    ref, _ = RL.load_ff_pwc(torch_correlation)
    ref.load_state_dict(model.state_dict(), strict=True)
    ref = ref.cuda().eval()
    im1, im2, m1, m2 = (t.cuda() / 255.0 for t in synthetic_pair(batch, hw[0], hw[1], seed=11))    # images in [0, 1]
    with torch.no_grad():
        want = ref(im1, im2, m1, m2, test_mode=True)
        got = model(im1, im2, m1, m2, test_mode=True)
        want_list = ref(im1, im2, m1, m2)
        got_list = model(im1, im2, m1, m2)
    assert got.shape == want.shape == (batch, 2, hw[0], hw[1])
    epe = torch.linalg.norm(got - want, dim=1)
    print(f"FF-PWC {hw} b{batch}: |flow| mean {float(want.abs().mean()):.2f} max {float(want.abs().max()):.2f}; "
          f"EPE mean {float(epe.mean()):.2e} max {float(epe.max()):.2e}")
    scale = max(1.0, float(want.abs().max()))
    assert float(want.abs().max()) > 0.5                      # the flow is not trivially zero
    assert float(epe.max()) <= 2e-4 * scale                   # random weights give large flows: relative to their size
    assert len(got_list) == len(want_list) == 5
    for a, r in zip(got_list, want_list):
        assert a.shape == r.shape and float((a - r).abs().max()) <= 5e-4 * max(1.0, float(r.abs().max()))
