"""Generate tests/golden/*.npz by running the UNMODIFIED reference on CPU.

TEST INFRASTRUCTURE ONLY.  Runs in the build container, where the reference is
mounted read-only at /root/reference; the GPU box has no such mount, so the
vectors are committed.  Usage::

    python oracle/make_golden.py            # corr fixtures (+ e2e if host model importable)

What is recorded
  corr_*.npz   inputs (fmap1, fmap2, coords_k) and the reference outputs of
               ``FF_RAFT_Core/corr.py`` CorrBlock: the 4 pyramid levels and the
               lookup result for each coords_k.
  ffraft_e2e.npz  inputs + ``FF_RAFT_FUSION`` test-mode outputs (flow_lo, flow_up)
               for weights filled by ``tests/weights.py`` (deterministic by key name,
               so no 30 MB state_dict needs to be committed).
"""
from __future__ import annotations

import argparse
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/core/models/ff-raft"
GOLD = os.path.join(ROOT, "tests", "golden")


def _ref_corr():
    sys.path.insert(0, REF)
    from FF_RAFT_Core.corr import CorrBlock  # noqa: E402  (the reference's own class)
    from FF_RAFT_Core.utils.utils import coords_grid  # noqa: E402
    return CorrBlock, coords_grid


def coords_cases(rng: np.random.Generator, grid: np.ndarray) -> dict[str, np.ndarray]:
    """Coordinate sets covering the edge cases SURVEY §8c T2 names."""
    b, _, h, w = grid.shape
    n = lambda s: rng.standard_normal(grid.shape).astype(np.float32) * np.float32(s)
    cases = {
        "grid": grid.copy(),                                   # exact integers (iteration 0)
        "half": grid + np.float32(0.5),                        # exact half-integers
        "s1": grid + n(1.0),
        "s3": grid + n(3.0),
        "s20": grid + n(20.0),                                 # ~40% out-of-bounds taps
        "far": grid + n(200.0),                                # almost everything outside
        "edge": grid * np.float32(0) + np.float32(-0.25),      # hugging the top-left corner
    }
    e2 = grid.copy()
    e2[:, 0] = np.float32(w - 1) + np.float32(0.75)
    e2[:, 1] = np.float32(h - 1) - np.float32(0.125)
    cases["edge_br"] = e2
    return {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in cases.items()}


def make_corr(name: str, b: int, d: int, h: int, w: int, seed: int, scale: float = 4.4, keep=None):
    CorrBlock, coords_grid = _ref_corr()
    rng = np.random.default_rng(seed)
    f1 = (rng.standard_normal((b, d, h, w)) * scale).astype(np.float32)
    f2 = (rng.standard_normal((b, d, h, w)) * scale).astype(np.float32)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        blk = CorrBlock(torch.from_numpy(f1), torch.from_numpy(f2), num_levels=4, radius=4)
        grid = coords_grid(b, h, w, device="cpu").numpy()
        cases = coords_cases(rng, grid)
        if keep is not None:
            cases = {k: v for k, v in cases.items() if k in keep}
        out = {"fmap1": f1, "fmap2": f2}
        for i, lvl in enumerate(blk.corr_pyramid):
            out[f"level{i}"] = lvl.numpy()[:, 0]
        for k, c in cases.items():
            out[f"coords_{k}"] = c
            out[f"lookup_{k}"] = blk(torch.from_numpy(c)).numpy()
    path = os.path.join(GOLD, f"corr_{name}.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items() if k.startswith(("level", "lookup_grid"))})


def make_e2e():
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, REF)
    from weights import fill_state_dict, synthetic_pair  # tests/weights.py
    from common import yaml_parser  # reference
    from FF_RAFT_Core.ff_raft import FF_RAFT_FUSION  # reference

    cfg = yaml_parser(os.path.join(REF, "config/experiment/ffraft_chairs_orb.yaml"))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = FF_RAFT_FUSION(use_fusion="parallel", fusion_channels=cfg.MODEL.FUSION_CHANNEL,
                               raft_small=False, dropout=0.0, alternate_corr=False,
                               abandon_fnet=False, fuse_cnet=True, cfg=cfg)
        sd = model.state_dict()
        fill_state_dict(sd, seed=1234)
        model.load_state_dict(sd, strict=True)
        model.eval()
        out = {}
        for tag, (b, hh, ww, iters) in {"a": (1, 128, 192, 12), "b": (2, 136, 160, 4)}.items():
            im1, im2, m1, m2 = synthetic_pair(b, hh, ww, seed=1234 + b)
            with torch.no_grad():
                lo, up = model(im1, im2, m1, m2, raft_iters=iters, test_mode=True)
            out[f"{tag}_shape"] = np.array([b, hh, ww, iters])
            out[f"{tag}_flow_lo"] = lo.numpy()
            out[f"{tag}_flow_up"] = up.numpy()
            print(tag, "flow_up mean |f| =", float(up.abs().mean()), "max", float(up.abs().max()))
    out["state_keys"] = np.array(sorted(sd.keys()))
    path = os.path.join(GOLD, "ffraft_e2e.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


# ---------------------------------------------------------------------------------------------------------------
# BASELINE shapes.  Inputs are pure functions of a seed (tests/weights.py: seeded_fmaps / seeded_coords /
# synthetic_pair / fill_state_dict), so the files hold only reference OUTPUTS, and those on a fixed subset:
#   fullsize_corr_*.npz   pyramid maps of `CorrBlock` for 24 queries, lookups for NQ evenly spaced queries
#   ffraft_e2e_full.npz  `FF_RAFT_FUSION` test-mode flow: the 1/8-resolution flow in full, flow_up on every 4th pixel
# ---------------------------------------------------------------------------------------------------------------
FULL_CORR = {  # name: (b, d, h, w, seed)
    "c1_46x62": (1, 256, 46, 62, 2101),       # BASELINE config 1 / 5 (368x496)
    "c2_47x156": (1, 256, 47, 156, 2102),     # BASELINE config 2 (376x1248)
}
FULL_COORDS = {"grid": (0.0, 0.0), "half": (0.0, 0.5), "s3": (3.0, 0.37), "s20": (20.0, 0.0)}   # name: (sigma, offset)
NQ = 192


def make_corr_full():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from weights import seeded_coords, seeded_fmaps

    CorrBlock, _ = _ref_corr()
    for name, (b, d, h, w, seed) in FULL_CORR.items():
        f1, f2 = seeded_fmaps(seed, b, d, h, w)
        n = h * w
        sel = np.unique(np.concatenate([np.linspace(0, b * n - 1, NQ).astype(np.int64), [0, w - 1, n - w, n - 1]]))
        sel_lv = np.unique(np.concatenate([sel[::10], [0, w - 1, n - w, n - 1]]))      # a subset of `queries`
        out = {"shape": np.array([b, d, h, w, seed]), "queries": sel, "queries_levels": sel_lv}
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            blk = CorrBlock(torch.from_numpy(f1), torch.from_numpy(f2), num_levels=4, radius=4)
            for i, lvl in enumerate(blk.corr_pyramid):
                out[f"level{i}"] = lvl.numpy()[sel_lv, 0]
            out["level0_absmax"] = np.float32(blk.corr_pyramid[0].abs().max())
            for k, (sigma, offset) in FULL_COORDS.items():
                c = seeded_coords(seed + 7, b, h, w, sigma, offset)
                res = blk(torch.from_numpy(c)).numpy().reshape(b, 324, n)            # [B, 324, N]
                out[f"lookup_{k}"] = np.ascontiguousarray(res.transpose(0, 2, 1).reshape(b * n, 324)[sel])
                out[f"lookup_{k}_absmax"] = np.float32(np.abs(res).max())
        path = os.path.join(GOLD, f"fullsize_corr_{name}.npz")
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path) >> 10, "KiB")


FULL_E2E = {  # tag: (b, H, W, iters, flow_gain name)
    "c1": (1, 368, 496, 12, "damped"),        # BASELINE config 1
    "c2": (1, 376, 1248, 12, "damped"),       # BASELINE config 2, one pair of the batch
    "c4": (1, 440, 1024, 32, "damped"),       # BASELINE config 4: Sintel 436x1024 padded (utils.py:9-16), 32 iterations
}
# Why no end-to-end EPE case with the "lively" flow head (2.5 px of motion per iteration): with random weights that
# refinement is chaotic.  Measured here on the reference itself (CPU, 368x496): rounding only the CorrBlock inputs to
# TF32 -- the reference's own GPU configuration -- moves its 12-iteration flow by 3e-3, 1e-2, 0.1, 0.24, 0.7, 1.5 ...
# 9.5 px (max EPE per iteration, x2.5 per iteration), the damped head by 3e-3 px in total.  A 0.01 px bar on that
# trajectory tests the chaos, not the kernels; make_trajectory() below pins the lively case where it is well-posed:
# the lookups along the REFERENCE's own coordinate trajectory.


def make_e2e_full():
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle.reference_loader import load_ff_raft
    from weights import DAMPED_GAIN, LIVELY_GAIN, fill_state_dict, synthetic_pair

    model, _, ns = load_ff_raft()
    raft = ns["raft"]
    base = raft.CorrBlock

    class Tf32Inputs(base):
        """The reference's own GPU arithmetic for the volume (ALLOW_TF32: operands rounded to a 10-bit mantissa), emulated
        on CPU by rounding the CorrBlock inputs: measures how far the reference drifts from ITSELF between its two
        configurations -- the yardstick for any tensor-core implementation of the volume."""

        def __init__(self, f1, f2, **kw):
            def rn(x):
                i = x.contiguous().view(torch.int32)
                return ((i + 0x1000) & ~0x1FFF).view(torch.float32)
            super().__init__(rn(f1.clone()), rn(f2.clone()), **kw)

    out = {}
    for tag, (b, hh, ww, iters, gain) in FULL_E2E.items():
        sd = model.state_dict()
        fill_state_dict(sd, seed=1234, flow_gain=LIVELY_GAIN if gain == "lively" else DAMPED_GAIN)
        model.load_state_dict(sd, strict=True)
        model.eval()
        im1, im2, m1, m2 = synthetic_pair(b, hh, ww, seed=4321)
        with torch.no_grad(), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            lo, up = model(im1, im2, m1, m2, raft_iters=iters, test_mode=True)
            raft.CorrBlock = Tf32Inputs
            try:
                lo_t, up_t = model(im1, im2, m1, m2, raft_iters=iters, test_mode=True)
            finally:
                raft.CorrBlock = base
        drift = torch.linalg.norm(up_t - up, dim=1)
        out[f"{tag}_shape"] = np.array([b, hh, ww, iters])
        out[f"{tag}_gain"] = np.array(gain)
        out[f"{tag}_flow_lo"] = lo.numpy()
        out[f"{tag}_flow_up_s4"] = np.ascontiguousarray(up.numpy()[:, :, ::4, ::4])
        out[f"{tag}_ref_tf32_drift"] = np.array([float(drift.mean()), float(drift.max())])
        print(tag, "flow_up mean |f| =", float(up.abs().mean()), "max", float(up.abs().max()),
              "| reference TF32-vs-fp32 drift: mean %.2e max %.2e px" % (float(drift.mean()), float(drift.max())))
    path = os.path.join(GOLD, "ffraft_e2e_full.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) >> 10, "KiB")


def make_trajectory():
    """Reference FF_RAFT_FUSION with the lively flow head at 368x496: the coordinates it hands to CorrBlock.__call__ at
    each of its 12 iterations (multi-pixel motion, windows far from the integer grid) and the lookup results for a fixed
    query subset.  Recorded by a subclass of the reference's CorrBlock swapped into its `raft` module namespace (the way
    raft.py:198 instantiates it by name); the reference code itself is untouched."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle.reference_loader import load_ff_raft
    from weights import LIVELY_GAIN, fill_state_dict, synthetic_pair

    model, _, ns = load_ff_raft()
    raft = ns["raft"]
    base = raft.CorrBlock
    rec = {"coords": [], "out": [], "absmax": []}
    b, hh, ww, iters = 1, 368, 496, 12
    n = (hh // 8) * (ww // 8)
    sel = np.unique(np.concatenate([np.linspace(0, n - 1, 60).astype(np.int64), [0, ww // 8 - 1, n - ww // 8, n - 1]]))

    class Recording(base):
        def __call__(self, coords):
            out = super().__call__(coords)
            rec["coords"].append(coords.detach().numpy().copy())
            rec["out"].append(out.detach().numpy().reshape(b, 324, n)[0].T[sel].copy())
            rec["absmax"].append(float(out.abs().max()))
            return out

    sd = model.state_dict()
    fill_state_dict(sd, seed=1234, flow_gain=LIVELY_GAIN)
    model.load_state_dict(sd, strict=True)
    model.eval()
    im1, im2, m1, m2 = synthetic_pair(b, hh, ww, seed=4321)
    raft.CorrBlock = Recording
    try:
        with torch.no_grad(), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            lo, up = model(im1, im2, m1, m2, raft_iters=iters, test_mode=True)
    finally:
        raft.CorrBlock = base
    out = {"shape": np.array([b, hh, ww, iters]), "queries": sel, "coords": np.stack(rec["coords"]),
           "lookups": np.stack(rec["out"]), "absmax": np.array(rec["absmax"]), "flow_lo": lo.numpy()}
    motion = np.abs(out["coords"][-1] - out["coords"][0])
    print("trajectory: final |coords - grid| mean", float(motion.mean()), "max", float(motion.max()), "low-res px")
    path = os.path.join(GOLD, "ffraft_trajectory_c1_lively.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) >> 10, "KiB")


# ---------------------------------------------------------------------------------------------------------------
# Training (BASELINE config 5).  T6 of SURVEY 8c: gradients of fmap1 / fmap2 through the reference CorrBlock's own
# autograd at the 368x496 feature-map shape, and one full MixLoss training step of FF_RAFT_FUSION (loss value,
# gradient norm of every parameter, three gradients in full).
# ---------------------------------------------------------------------------------------------------------------
def make_corr_grad():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from weights import seeded_coords, seeded_fmaps

    CorrBlock, _ = _ref_corr()
    b, d, h, w, seed = 1, 256, 46, 62, 2201
    f1, f2 = seeded_fmaps(seed, b, d, h, w)
    t1, t2 = torch.from_numpy(f1).requires_grad_(True), torch.from_numpy(f2).requires_grad_(True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        blk = CorrBlock(t1, t2, num_levels=4, radius=4)
        loss = 0.0
        for k, (sigma, offset) in enumerate([(0.0, 0.0), (2.0, 0.3), (6.0, 0.0)]):      # 3 lookups share one pyramid
            c = torch.from_numpy(seeded_coords(seed + 10 + k, b, h, w, sigma, offset))
            g = torch.from_numpy(np.random.RandomState(seed + 20 + k).standard_normal((b, 324, h, w)).astype(np.float32))
            loss = loss + (blk(c) * g).sum()
        loss.backward()
    out = {"shape": np.array([b, d, h, w, seed]), "loss": np.float64(loss.item()),
           "gfmap1_s16": t1.grad.numpy()[:, ::16].copy(), "gfmap2_s16": t2.grad.numpy()[:, ::16].copy(),
           "gfmap1_norm": np.float64(t1.grad.double().norm()), "gfmap2_norm": np.float64(t2.grad.double().norm())}
    path = os.path.join(GOLD, "fullsize_grad_c5_46x62.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) >> 10, "KiB")


TRAIN_CASES = {"small": (2, 128, 160, 3), "c5": (1, 368, 496, 12)}     # tag: (b, H, W, iters)
TRAIN_FULL_GRADS = ["flow_net.fnet.conv2.bias", "flow_net.update_block.flow_head.conv2.weight",
                    "flow_net.fnet.fusion5.mask2img.conv.bias", "flow_net.cnet.layer1.0.norm1.weight"]


def make_train():
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle.reference_loader import load_ff_raft
    from weights import fill_state_dict, train_inputs

    model, cfg, ns = load_ff_raft()
    sys.path.insert(0, ns["root"])
    from losses import build_losses  # the reference's losses/__init__.py

    loss_fn = build_losses(cfg.TRAIN.LOSS_TYPE, gamma=cfg.TRAIN.LOSS_GAMMA, max_flow=cfg.TRAIN.MAX_FLOW,
                           kernel_size=cfg.TRAIN.LOSS_KERNEL_SIZE, sigma=cfg.TRAIN.LOSS_SIGMA, lamda=cfg.TRAIN.LOSS_LAMDA)
    out = {}
    for tag, (b, hh, ww, iters) in TRAIN_CASES.items():
        sd = model.state_dict()
        fill_state_dict(sd, seed=1234)
        model.load_state_dict(sd, strict=True)
        model.train()                                            # chairs stage: BatchNorm in training mode (train.py:192)
        model.zero_grad(set_to_none=True)
        im1, im2, flow, m1, m2, valid = train_inputs(b, hh, ww, 777)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            preds = model(im1, im2, m1, m2, raft_iters=iters)
            loss, metrics = loss_fn(preds, flow, valid, m1)
            loss.backward()
        names = [k for k, p in model.named_parameters() if p.grad is not None]
        out[f"{tag}_shape"] = np.array([b, hh, ww, iters])
        out[f"{tag}_loss"] = np.float64(loss.item())
        out[f"{tag}_epe"] = np.float64(metrics["epe"])
        out[f"{tag}_param_names"] = np.array(names)
        out[f"{tag}_grad_norms"] = np.array([float(dict(model.named_parameters())[k].grad.double().norm()) for k in names])
        for k in TRAIN_FULL_GRADS:
            out[f"{tag}_grad::{k}"] = dict(model.named_parameters())[k].grad.numpy().copy()
        print(tag, "loss", loss.item(), "epe", metrics["epe"], "params with grad", len(names))
    path = os.path.join(GOLD, "ffraft_train_step.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) >> 10, "KiB")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", choices=["corr", "e2e", "corr_full", "e2e_full", "trajectory", "corr_grad", "train"], default=None)
    a = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(8)
    if a.only in (None, "corr"):
        make_corr("b1_d32_16x24", b=1, d=32, h=16, w=24, seed=11)
        make_corr("b2_d16_17x21", b=2, d=16, h=17, w=21, seed=12, keep=("grid", "s3", "s20"))      # odd dims: floor pooling
        make_corr("b1_d256_16x16", b=1, d=256, h=16, w=16, seed=13, keep=("grid", "s1"))    # the real D
    if a.only in (None, "e2e"):
        make_e2e()
    if a.only in (None, "corr_full"):
        make_corr_full()
    if a.only in (None, "e2e_full"):
        make_e2e_full()
    if a.only in (None, "trajectory"):
        make_trajectory()
    if a.only in (None, "corr_grad"):
        make_corr_grad()
    if a.only in (None, "train"):
        make_train()
