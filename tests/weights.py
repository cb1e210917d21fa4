"""Deterministic weights and synthetic inputs shared by the golden generator
(oracle/make_golden.py, runs the reference) and the tests / bench (run this repo's host model).

Weights are a pure function of (seed, state_dict key, shape), so the reference model and the
re-written host model get bit-identical parameters without committing a 30 MB checkpoint.
"""
from __future__ import annotations

import zlib

import numpy as np
import torch


DAMPED_GAIN = 0.004    # sub-pixel updates per iteration, like a converged trained network
LIVELY_GAIN = 0.05     # ~0.3 low-res px (2.5 full-res px) per iteration: multi-pixel motion, windows far from the grid


PWC_GAINS = {"netSix.0.weight": 0.1, "netMain.12.weight": 0.1}    # FF-PWC flow heads: keep the coarse-to-fine flow O(1 px)


def fill_state_dict(sd: dict, seed: int = 1234, flow_gain: float = DAMPED_GAIN, gains: dict = None) -> dict:
    """In-place, key-wise deterministic fill (kaiming-like scale so activations stay O(1)).

    `flow_gain` scales the last conv of the flow head: DAMPED_GAIN keeps the random-init refinement contractive
    (every iteration moves the flow by ~0.2 px at full resolution); LIVELY_GAIN lets it move ~2.5 px per iteration
    (27 px mean / 90 px max after 12 iterations at 368x496), so the lookups leave the integer grid by several cells
    and the end-to-end comparison becomes sensitive to the correlation values themselves."""
    for key in sorted(sd.keys()):
        t = sd[key]
        rng = np.random.RandomState((seed * 1000003 + zlib.crc32(key.encode())) % (2 ** 31 - 1))
        if key.endswith("num_batches_tracked"):
            t.fill_(1)
            continue
        shape = tuple(t.shape)
        if key.endswith("running_var"):
            v = 1.0 + 0.1 * rng.rand(*shape)
        elif key.endswith("running_mean"):
            v = 0.05 * rng.randn(*shape)
        elif t.dim() == 4:  # conv weight [out, in, kh, kw]: fan_out kaiming like the reference init
            fan_out = shape[0] * shape[2] * shape[3]
            v = rng.randn(*shape) * np.sqrt(2.0 / fan_out)
            if key.endswith("flow_head.conv2.weight"):
                v = v * flow_gain
            for suffix, factor in (gains or {}).items():
                if key.endswith(suffix):
                    v = v * factor
        elif key.endswith("weight"):  # norm scale
            v = 1.0 + 0.05 * rng.randn(*shape)
        else:  # biases
            v = 0.02 * rng.randn(*shape)
        t.copy_(torch.from_numpy(np.asarray(v, dtype=np.float32)).reshape(shape))
    return sd


def synthetic_pair(batch: int, height: int, width: int, seed: int = 1234, keypoints: int = 500, border: int = 31):
    """SURVEY 8d synthetic inputs: image1 ~ U[0,255), image2 = image1 shifted by (3,5) px + N(0,2),
    ORB-style point mask (255 at `keypoints` random pixels >= `border` px from the edge)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    im1 = torch.rand(batch, 3, height, width, generator=g) * 255.0
    im2 = torch.roll(im1, shifts=(3, 5), dims=(2, 3)) + torch.randn(batch, 3, height, width, generator=g) * 2.0
    im2 = im2.clamp_(0.0, 255.0)
    bd = min(border, height // 4, width // 4)
    mask = torch.zeros(batch, 1, height, width)
    ys = torch.randint(bd, height - bd, (batch, keypoints), generator=g)
    xs = torch.randint(bd, width - bd, (batch, keypoints), generator=g)
    for b in range(batch):
        mask[b, 0, ys[b], xs[b]] = 255.0
    return im1, im2, mask, mask.clone()


def seeded_fmaps(seed: int, b: int, d: int, h: int, w: int, scale: float = 4.4):
    """Feature maps as a pure function of the seed (legacy RandomState: bit-stable across numpy versions), so that
    full-size golden files need not store their 7.5 MB inputs."""
    rng = np.random.RandomState(seed)
    f1 = (rng.standard_normal((b, d, h, w)) * scale).astype(np.float32)
    f2 = (rng.standard_normal((b, d, h, w)) * scale).astype(np.float32)
    return f1, f2


def seeded_coords(seed: int, b: int, h: int, w: int, sigma: float, offset: float = 0.0):
    """coords_grid (x, y order) + offset + sigma * N(0,1), a pure function of the seed."""
    rng = np.random.RandomState(seed)
    ys, xs = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    grid = np.repeat(np.stack([xs, ys], 0).astype(np.float32)[None], b, axis=0)
    noise = rng.standard_normal(grid.shape).astype(np.float32) * np.float32(sigma) if sigma else np.float32(0)
    return np.ascontiguousarray(grid + np.float32(offset) + noise, dtype=np.float32)


def train_inputs(b: int, hh: int, ww: int, seed: int):
    """Inputs of one training step (image1, image2, flow_gt, mask1, mask2, valid) as a pure function of the seed; shared
    by oracle/make_golden.py (reference step), tests/test_training.py and bench.py --config 5."""
    im1, im2, m1, m2 = synthetic_pair(b, hh, ww, seed=seed)
    rng = np.random.RandomState(seed + 1)
    flow = torch.from_numpy((rng.standard_normal((b, 2, hh, ww)) * 3.0).astype(np.float32))
    valid = torch.from_numpy((rng.uniform(size=(b, hh, ww)) > 0.1).astype(np.float32))
    return im1, im2, flow, m1, m2, valid
