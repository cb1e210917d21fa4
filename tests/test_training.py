"""Training path (BASELINE config 5): MixLoss, one training step of the host model, gradient parity with the reference.

CPU tests: the host model in train mode + a differentiable CorrBlock made of the reference's own ATen ops
(oracle/corr_torch_cpu.py) against tests/golden/ffraft_train_step.npz = loss and per-parameter gradients of the
UNMODIFIED FF_RAFT_FUSION + the reference's MixLoss (oracle/make_golden.py --only train).
GPU tests (-m gpu): the same step with the B200 CorrBlock (native forward AND backward kernels), and the gradients of
fmap1 / fmap2 through the correlation path against the reference's autograd at the 368x496 feature-map shape.
"""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
from weights import fill_state_dict, seeded_coords, seeded_fmaps, train_inputs  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def host_model(device="cpu"):
    from focusflow_official_b200.host import FocusRAFT

    model = FocusRAFT()
    sd = model.state_dict()
    fill_state_dict(sd, seed=1234)
    model.load_state_dict(sd, strict=True)
    return model.to(device)


def check_step_against_golden(model, tag, device, norm_tol, loss_tol):
    from focusflow_official_b200.host import build_losses

    g = np.load(os.path.join(GOLD, "ffraft_train_step.npz"))
    b, hh, ww, iters = [int(v) for v in g[f"{tag}_shape"]]
    loss_fn = build_losses("MixLoss", gamma=0.8, max_flow=400, kernel_size=1, sigma=0.01, lamda=1)
    model.train()
    model.zero_grad(set_to_none=True)
    im1, im2, flow, m1, m2, valid = (x.to(device) for x in train_inputs(b, hh, ww, 777))
    preds = model(im1, im2, m1, m2, raft_iters=iters)
    assert len(preds) == iters
    loss, metrics = loss_fn(preds, flow, valid, m1)
    loss.backward()
    assert abs(loss.item() - float(g[f"{tag}_loss"])) <= loss_tol * abs(float(g[f"{tag}_loss"]))
    assert abs(metrics["epe"] - float(g[f"{tag}_epe"])) <= loss_tol * float(g[f"{tag}_epe"])
    params = dict(model.named_parameters())
    names = [str(k) for k in g[f"{tag}_param_names"]]
    assert sorted(names) == sorted(k for k, p in params.items() if p.grad is not None) and len(names) == 254
    ref_norms = g[f"{tag}_grad_norms"]
    got = np.array([float(params[k].grad.double().norm()) for k in names])
    rel = np.abs(got - ref_norms) / np.maximum(ref_norms, 1e-3 * ref_norms.max())
    assert rel.max() <= norm_tol, (names[int(rel.argmax())], float(rel.max()))
    for key in [k for k in g.files if k.startswith(f"{tag}_grad::")]:
        ref = g[key]
        have = params[key.split("::", 1)[1]].grad.detach().cpu().numpy()
        rel = np.linalg.norm(have - ref) / max(np.linalg.norm(ref), 1e-30)
        assert rel <= 4 * norm_tol, (key, float(rel))
    return float(loss.item())


def test_losses_match_their_definition():
    """losses.py:18-138 on random tensors: the three losses are weighted sums of the same two terms."""
    from focusflow_official_b200.host import CPCL, EPELoss, MixLoss, build_losses
    from focusflow_official_b200.host.losses import get_kernel

    torch.manual_seed(0)
    b, h, w, n = 2, 12, 16, 4
    preds = [torch.randn(b, 2, h, w) * 2 for _ in range(n)]
    gt = torch.randn(b, 2, h, w) * 2
    gt[0, :, 0, 0] = 500.0                                   # beyond max_flow: excluded
    valid = (torch.rand(b, h, w) > 0.2).float()
    mask = (torch.rand(b, 1, h, w) > 0.9).float() * 255
    ok = ((valid >= 0.5) & (gt.pow(2).sum(1).sqrt() < 400))[:, None].float()
    m = (mask > 0).float()
    dense = sum(0.8 ** (n - i - 1) * (ok * (p - gt).abs()).mean() for i, p in enumerate(preds))
    point = sum(0.8 ** (n - i - 1) * (ok * m * (p - gt).abs()).sum() / m.sum() for i, p in enumerate(preds))
    k1 = dict(kernel_size=1, sigma=0.01)
    assert torch.allclose(EPELoss(0.8, 400)(preds, gt, valid)[0], dense, rtol=1e-6)
    assert torch.allclose(CPCL(0.8, 400, **k1)(preds, gt, valid, mask)[0], point, rtol=1e-6)
    loss, metrics = MixLoss(0.8, 400, lamda=0.7, **k1)(preds, gt, valid, mask)
    assert torch.allclose(loss, dense + 0.7 * point, rtol=1e-6)
    epe = (preds[-1] - gt).pow(2).sum(1).sqrt()[ok[:, 0] > 0].mean()
    assert abs(metrics["epe"] - float(epe)) < 1e-6 and abs(metrics["loss"] - float(loss)) < 1e-6
    # a real Gaussian spreads the key points; the kernel sums to one
    k = get_kernel(5, 1.7)
    assert k.shape == (1, 1, 5, 5) and abs(float(k.sum()) - 1) < 1e-6 and float(k[0, 0, 2, 2]) == float(k.max())
    assert isinstance(build_losses("MixLoss"), MixLoss) and isinstance(build_losses("CPCL"), CPCL)
    with pytest.raises(ValueError):
        build_losses("L2")


def test_host_training_step_matches_the_reference_on_cpu():
    """Train-mode host model (batch-norm statistics, all 12 predictions, convex upsampling) + MixLoss + autograd of the
    reference's own correlation ops == FF_RAFT_FUSION + reference MixLoss: loss, EPE and all 254 parameter gradients."""
    from corr_torch_cpu import TorchCorrBlock

    torch.set_num_threads(8)
    model = host_model()
    model.flow_net.corr_block = TorchCorrBlock
    check_step_against_golden(model, "small", "cpu", norm_tol=2e-3, loss_tol=1e-5)


def test_train_step_object_updates_the_weights():
    """TrainStep = train.py:296-328: zero_grad, forward, MixLoss, backward, clip 1.0, AdamW, OneCycleLR."""
    from corr_torch_cpu import TorchCorrBlock
    from focusflow_official_b200.host import TrainStep

    model = host_model()
    model.flow_net.corr_block = TorchCorrBlock
    step = TrainStep(model, iters=2, num_steps=100)
    before = model.flow_net.update_block.flow_head.conv2.weight.detach().clone()
    batch = train_inputs(1, 128, 128, 5)     # 16x16 feature map: the coarsest level is 2x2 (1x1 is degenerate)
    loss0, metrics = step(*batch)
    assert torch.isfinite(loss0) and set(metrics) == {"epe", "loss"} and float(step.grad_norm) > 0
    assert not torch.equal(before, model.flow_net.update_block.flow_head.conv2.weight)
    assert step.scheduler.last_epoch == 1
    lr0 = step.optimizer.param_groups[0]["lr"]
    step(*batch)
    assert step.optimizer.param_groups[0]["lr"] > lr0        # warm-up phase of the one-cycle schedule


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("precision,tol", [("fp32", 2e-4), ("fp16", 2e-3), ("tf32", 2e-3)])
def test_corr_gradients_match_the_reference_autograd_at_config5_shape(precision, tol):
    """d loss / d fmap1, d fmap2 through build + 3 lookups, against the reference CorrBlock's own autograd
    (CPU fp32) at D=256, 46x62.  fp32 = exact CUDA-core path; fp16 / tf32 = tensor-core forward, tf32 backward GEMMs."""
    import focusflow_official_b200 as ff

    g = np.load(os.path.join(GOLD, "fullsize_grad_c5_46x62.npz"))
    b, d, h, w, seed = [int(v) for v in g["shape"]]
    f1, f2 = seeded_fmaps(seed, b, d, h, w)
    t1 = torch.from_numpy(f1).cuda().requires_grad_(True)
    t2 = torch.from_numpy(f2).cuda().requires_grad_(True)
    for channels_last in (False, True):
        t1.grad = t2.grad = None
        blk = ff.CorrBlock(t1, t2, num_levels=4, radius=4, precision=precision, sampler="cpu", channels_last=channels_last)
        loss = 0.0
        for k, (sigma, offset) in enumerate([(0.0, 0.0), (2.0, 0.3), (6.0, 0.0)]):
            c = torch.from_numpy(seeded_coords(seed + 10 + k, b, h, w, sigma, offset)).cuda()
            go = torch.from_numpy(np.random.RandomState(seed + 20 + k).standard_normal((b, 324, h, w)).astype(np.float32)).cuda()
            loss = loss + (blk(c) * go).sum()
        loss.backward()
        assert abs(loss.item() - float(g["loss"])) <= max(tol, 1e-3) * 5e5     # |loss| ~ sum of 3e5 products of O(20)
        for name, tt in (("gfmap1", t1), ("gfmap2", t2)):
            got = tt.grad[:, ::16].cpu().numpy()
            ref = g[f"{name}_s16"]
            rel = np.linalg.norm(got - ref) / np.linalg.norm(ref)
            assert rel <= tol, (precision, name, channels_last, rel)
            assert abs(float(tt.grad.double().norm()) - float(g[f"{name}_norm"])) <= tol * float(g[f"{name}_norm"])


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["small", "c5"])
def test_training_step_with_native_kernels_matches_the_reference(tag):
    """One MixLoss step at 128x160 x3 and at the config-5 shape (368x496, 12 iterations): host model on the GPU with the
    B200 CorrBlock (fused tcgen05 build, tiled lookups, native lookup / pyramid / volume backward) against the reference
    step recorded on CPU.  fp32 convolutions so that the comparison isolates the correlation path."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    model = host_model("cuda:0")
    model.flow_net.corr_sampler = "cpu"
    for prec, tol in (("fp32", 5e-3), ("fp16", 2e-2)):
        model.flow_net.corr_precision = prec
        check_step_against_golden(model, tag, "cuda:0", norm_tol=tol, loss_tol=1e-3)


# ------------------------------------------------------------------------------------------------ DDP under gloo, world 2
def _ddp_worker(rank, world, port, q):
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    from corr_torch_cpu import TorchCorrBlock
    from focusflow_official_b200.host import TrainStep

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    model = host_model()
    model.flow_net.corr_block = TorchCorrBlock
    ddp = DDP(model)                                           # what parallel_model() returns on a GPU rank (common.py:45-50)
    step = TrainStep(ddp, world_size=world, iters=2, num_steps=100)
    batch = train_inputs(1, 128, 128, 40 + rank)               # every rank its own pair
    loss, _ = step(*batch)
    loss2, _ = step(*batch)                                    # every rank takes the SAME number of steps (bench.py lesson)
    w = model.flow_net.update_block.flow_head.conv2.weight.detach()
    gathered = [torch.empty_like(w) for _ in range(world)]
    dist.all_gather(gathered, w.contiguous())
    q.put((rank, float(loss), float(loss2), bool(torch.equal(gathered[0], gathered[1])), float(step.grad_norm)))
    dist.barrier()
    dist.destroy_process_group()


def test_training_step_under_ddp_gloo_world2():
    """The N > 1 training path on CPU: stock DDP (gloo) around the host model, `loss *= world_size` (train.py:313-314),
    gradients all-reduced during backward, identical weights on both ranks after the optimizer steps."""
    import socket

    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[0] for r in res] == [0, 1]
    assert all(np.isfinite(r[1]) and np.isfinite(r[2]) for r in res)
    assert all(r[3] for r in res), "ranks diverged: the gradients were not all-reduced"
    assert res[0][1] != res[1][1]                              # different data per rank ...
    assert abs(res[0][4] - res[1][4]) < 1e-6 * max(1.0, res[0][4])   # ... but one (averaged x world) gradient
