"""Pins the CPU oracle against vectors produced by the UNMODIFIED reference
(oracle/make_golden.py ran /root/reference's CorrBlock on CPU).  CPU-only."""
import glob
import os

import numpy as np
import pytest

from oracle import corr_oracle as co
from oracle import pwc_oracle as po

FILES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "corr_*.npz")))


def _cases(g):
    return [k[len("coords_"):] for k in g.files if k.startswith("coords_")]


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(p) for p in FILES])
def test_volume_and_pyramid_match_reference(path):
    g = np.load(path)
    blk = co.CorrBlock(g["fmap1"], g["fmap2"], num_levels=4, radius=4)
    for i in range(4):
        ref = g[f"level{i}"]
        got = blk.corr_pyramid[i]
        assert got.shape == ref.shape
        # sgemm order may differ between BLAS builds; 1e-6 is ~4 ulp of the largest entry
        assert np.abs(got - ref).max() <= 1e-6 * np.abs(ref).max()
    # given the reference's own level 0 the pooled levels are bit-exact
    pyr = co.pyramid(g["level0"], 4)
    for i in range(1, 4):
        assert np.array_equal(pyr[i], g[f"level{i}"])


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(p) for p in FILES])
def test_lookup_matches_reference(path):
    g = np.load(path)
    pyr = [g[f"level{i}"] for i in range(4)]
    assert len(_cases(g)) >= 2
    for k in _cases(g):
        ref = g[f"lookup_{k}"]
        got = co.lookup(pyr, g[f"coords_{k}"], radius=4)
        assert got.shape == ref.shape and got.dtype == np.float32
        scale = max(float(np.abs(ref).max()), 1e-20)
        assert np.abs(got - ref).max() <= 1e-6 * scale, k


def test_lookup_channel_order_is_x_major():
    """corr.py:37-43: channel a*9+b samples (x + a - 4, y + b - 4)."""
    h, w = 12, 14
    q = h * w
    lvl = np.zeros((q, h, w), np.float32)
    lvl[:, 5, 9] = 1.0  # a delta at (y=5, x=9) in every query's map
    coords = np.zeros((1, 2, h, w), np.float32)
    coords[:, 0] = 7.0  # x
    coords[:, 1] = 6.0  # y
    out = co.lookup_level(lvl, coords[0, 0].ravel(), coords[0, 1].ravel(), 4)
    a, b = 9 - 7 + 4, 5 - 6 + 4
    hit = np.zeros(81, np.float32)
    hit[a * 9 + b] = 1.0
    assert np.allclose(out[0], hit, atol=1e-5)


def test_coords_grid_layout():
    g = co.coords_grid(2, 3, 5)
    assert g.shape == (2, 2, 3, 5)
    assert g[1, 0, 2, 4] == 4 and g[1, 1, 2, 4] == 2  # ch0 = x, ch1 = y


def test_pool_floor_and_order():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((3, 5, 7)).astype(np.float32) * 100
    y = co.pool2x2(x)
    assert y.shape == (3, 2, 3)
    ref = (((x[:, 0, 0] + x[:, 0, 1]) + x[:, 1, 0]) + x[:, 1, 1]) / np.float32(4)
    assert np.array_equal(y[:, 0, 0], ref)


def test_pwc_c_restatement_matches_closed_form():
    rng = np.random.default_rng(1)
    for (b, c, h, w) in [(2, 40, 9, 13), (1, 196, 7, 16), (1, 3, 5, 5)]:
        one = rng.standard_normal((b, c, h, w)).astype(np.float32)
        two = rng.standard_normal((b, c, h, w)).astype(np.float32)
        a = po.forward_c(one, two)
        n = po.forward_np(one, two)
        assert a.shape == (b, 81, h, w)
        assert np.abs(a - n).max() <= 2e-6 * max(1.0, np.abs(n).max())
        g = rng.standard_normal((b, 81, h, w)).astype(np.float32)
        c1, c2 = po.backward_c(one, two, g)
        n1, n2 = po.backward_np(one, two, g)
        assert np.abs(c1 - n1).max() <= 1e-5 and np.abs(c2 - n2).max() <= 1e-5


PWC_GOLD = os.path.join(os.path.dirname(__file__), "golden", "pwc_ref_cuda.npz")


def _pwc_inputs(idx, shape):
    """Same seeded inputs as oracle/make_golden_pwc.py."""
    rng = np.random.default_rng(1000 + idx)
    b, c, h, w = shape
    return (rng.standard_normal(shape).astype(np.float32), rng.standard_normal(shape).astype(np.float32),
            rng.standard_normal((b, 81, h, w)).astype(np.float32))


def test_pwc_oracle_matches_the_reference_cuda_kernels():
    """tests/golden/pwc_ref_cuda.npz holds the outputs of the REFERENCE's own CUDA kernels (correlation.py strings,
    specialised by its own cupy_kernel() and compiled with nvcc by oracle/build_pwc_ref_cuda.py, run on a B200).
    Both restatements must reproduce them: the C one keeps the kernel's 32-lane partial-sum order."""
    g = np.load(PWC_GOLD)
    idxs = sorted(int(k.split("_")[1]) for k in g.files if k.startswith("shape_"))
    assert idxs, "empty fixture"
    for idx in idxs:
        shape = tuple(int(v) for v in g[f"shape_{idx}"])
        one, two, gout = _pwc_inputs(idx, shape)
        ref, r1, r2 = g[f"out_{idx}"], g[f"gone_{idx}"], g[f"gtwo_{idx}"]
        for name, fwd, bwd, tol in (("c", po.forward_c, po.backward_c, 2e-7), ("numpy", po.forward_np, po.backward_np, 2e-6)):
            out = fwd(one, two)
            g1, g2 = bwd(one, two, gout)
            assert np.abs(out - ref).max() <= tol * max(1.0, np.abs(ref).max()), (idx, name, np.abs(out - ref).max())
            assert np.abs(g1 - r1).max() <= 10 * tol * max(1.0, np.abs(r1).max()), (idx, name, np.abs(g1 - r1).max())
            assert np.abs(g2 - r2).max() <= 10 * tol * max(1.0, np.abs(r2).max()), (idx, name, np.abs(g2 - r2).max())


def test_pwc_channel_order():
    """correlation.py:71-72: ch%9-4 shifts x, ch//9-4 shifts y."""
    one = np.zeros((1, 1, 9, 9), np.float32)
    two = np.zeros((1, 1, 9, 9), np.float32)
    one[0, 0, 4, 4] = 1.0
    two[0, 0, 6, 3] = 1.0  # dy = +2, dx = -1
    out = po.forward_c(one, two)
    ch = (2 + 4) * 9 + (-1 + 4)
    assert out[0, ch, 4, 4] == 1.0 and np.count_nonzero(out) == 1


def test_lookup_backward_is_adjoint():
    rng = np.random.default_rng(3)
    q, h, w = 6, 9, 11
    lvl = rng.standard_normal((q, h, w)).astype(np.float32)
    cx = (rng.random(q) * (w + 6) - 3).astype(np.float32)
    cy = (rng.random(q) * (h + 6) - 3).astype(np.float32)
    g = rng.standard_normal((q, 81)).astype(np.float32)
    fwd = co.lookup_level(lvl, cx, cy, 4)
    bwd = co.lookup_level_backward(g, lvl.shape, cx, cy, 4)
    lhs = float((fwd.astype(np.float64) * g).sum())
    rhs = float((bwd.astype(np.float64) * lvl).sum())
    assert abs(lhs - rhs) <= 1e-4 * max(1.0, abs(lhs))


# ---------------------------------------------------------------- BASELINE shapes (368x496 and 376x1248 feature maps)
FULL = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "fullsize_corr_*.npz")))


@pytest.mark.parametrize("path", FULL, ids=[os.path.basename(p) for p in FULL])
def test_oracle_matches_reference_at_baseline_shapes(path):
    """fullsize_corr_*.npz: reference CorrBlock outputs at D=256, 46x62 / 47x156 for a fixed query subset; the inputs
    are regenerated from the seed (tests/weights.py), so this also pins the seeded generators."""
    from weights import seeded_coords, seeded_fmaps

    g = np.load(path)
    b, d, h, w, seed = [int(v) for v in g["shape"]]
    assert len(FULL) == 2 and b == 1 and d == 256
    f1, f2 = seeded_fmaps(seed, b, d, h, w)
    n = h * w
    ql = g["queries_levels"]
    # volume rows of the selected queries, then their pyramids: the oracle's arithmetic on a subset
    rows = (f1.reshape(d, n)[:, ql].T @ f2.reshape(d, n)) / np.sqrt(np.float32(d))
    pyr = co.pyramid(rows.astype(np.float32).reshape(len(ql), h, w), 4)
    for i in range(4):
        ref = g[f"level{i}"]
        assert pyr[i].shape == ref.shape
        assert np.abs(pyr[i] - ref).max() <= 2e-6 * np.abs(g["level0"]).max(), i
    own = co.pyramid(g["level0"], 4)
    assert all(np.array_equal(own[i], g[f"level{i}"]) for i in range(1, 4))
    # lookups: queries present in both subsets, on the reference's own pyramid rows
    q = g["queries"]
    both, ia, ib = np.intersect1d(ql, q, return_indices=True)
    assert len(both) >= 20
    for k, (sigma, offset) in {"grid": (0.0, 0.0), "half": (0.0, 0.5), "s3": (3.0, 0.37), "s20": (20.0, 0.0)}.items():
        c = seeded_coords(seed + 7, b, h, w, sigma, offset).reshape(2, n)[:, both]
        got = np.concatenate([co.lookup_level(g[f"level{i}"][ia], c[0] / np.float32(2 ** i), c[1] / np.float32(2 ** i), 4)
                              for i in range(4)], axis=1)
        ref = g[f"lookup_{k}"][ib]
        assert np.abs(got - ref).max() <= 1e-6 * float(g[f"lookup_{k}_absmax"]), k
