"""GPU parity tests (run on the B200 box: ``pytest -m gpu``).

Every test drives the CUDA kernels through the package's public surface, i.e. through the
C ABI of libffcorr.so, and compares with the CPU oracle (``oracle/``) or with the golden
vectors the unmodified reference produced (``tests/golden``).  Tolerances are the ones
BASELINE.json states: volume <= 1e-3 relative, lookups <= 1e-5 (relative to max |ref|),
pyramid bit-exact given the same level 0.
"""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import corr_oracle as co
from oracle import pwc_oracle as po

pytestmark = pytest.mark.gpu

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "corr_*.npz")))
DEV = "cuda:0"


def ff():
    import focusflow_official_b200 as m

    return m


@pytest.fixture(autouse=True)
def _cpu_sampler_default():
    """The oracle and the golden vectors follow the reference's CPU run (ATen divides by W-1), so inside this module
    blocks and calls that do not name a sampler use "cpu"; the package default is "cuda" (tested in test_host_cpu.py,
    and against ATen's CUDA grid_sample in test_lookup_vs_the_reference_formula_on_torch_cuda)."""
    m = ff()
    before = m.get_sampler_semantics()
    m.set_sampler_semantics("cpu")
    yield
    m.set_sampler_semantics(before)


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def rel_fro(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def max_rel(a, b):
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() / max(np.abs(b).max(), 1e-30))


# ---------------------------------------------------------------- golden vectors
@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_golden_pyramid_bit_exact(path):
    g = np.load(path)
    from focusflow_official_b200 import _lib

    lv = [t(g["level0"][:, None])] + [torch.empty((g[f"level{i}"].shape[0], 1) + g[f"level{i}"].shape[1:], device=DEV)
                                       for i in range(1, 4)]
    q, _, h, w = lv[0].shape
    _lib.check(_lib.lib().ffcorr_pyramid_f32(_lib.ptr_array(lv), 4, q, h, w, _lib.current_stream()), "pyramid")
    for i in range(1, 4):
        assert np.array_equal(lv[i][:, 0].cpu().numpy(), g[f"level{i}"]), f"level {i}"


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_golden_lookup(path):
    g = np.load(path)
    levels = [t(g[f"level{i}"][:, None]) for i in range(4)]
    for k in [k[len("coords_"):] for k in g.files if k.startswith("coords_")]:
        out = ff().lookup(levels, t(g[f"coords_{k}"]), 4).cpu().numpy()
        ref = g[f"lookup_{k}"]
        assert out.shape == ref.shape
        assert max_rel(out, ref) <= 1e-5, (k, max_rel(out, ref))


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
@pytest.mark.parametrize("prec,tol", [("fp32", 2e-6), ("bf16x3", 3e-5), ("fp16", 1e-3), ("tf32", 1e-3)])
def test_golden_volume(path, prec, tol):
    g = np.load(path)
    blk = ff().CorrBlock(t(g["fmap1"]), t(g["fmap2"]), num_levels=4, radius=4, precision=prec)
    for i in range(4):
        got = blk.corr_pyramid[i][:, 0].cpu().numpy()
        assert rel_fro(got, g[f"level{i}"]) <= tol, (prec, i, rel_fro(got, g[f"level{i}"]))
    assert blk.corr_pyramid[0].shape == (g["level0"].shape[0], 1) + g["level0"].shape[1:]


# ---------------------------------------------------------------- volume vs oracle
@pytest.mark.parametrize("shape", [(1, 256, 46, 62), (2, 256, 17, 21), (1, 64, 16, 130), (3, 40, 9, 13), (1, 256, 24, 32)])
@pytest.mark.parametrize("prec,tol", [("fp16", 1e-3), ("tf32", 1e-3), ("bf16x3", 3e-5), ("fp32", 2e-6)])
def test_volume_vs_oracle(shape, prec, tol):
    b, d, h, w = shape
    rng = np.random.default_rng(1234)
    f1 = (rng.standard_normal(shape) * 4.4).astype(np.float32)
    f2 = (rng.standard_normal(shape) * 4.4).astype(np.float32)
    ref = co.volume_f64(f1, f2)
    got = ff().correlation_volume(t(f1), t(f2), precision=prec).cpu().numpy().reshape(ref.shape)
    err = rel_fro(got, ref)
    assert err <= tol, (shape, prec, err)
    # no element may be wildly off (catches tile / swizzle / tail mistakes that a norm hides)
    assert np.abs(got - ref).max() <= 50 * tol * np.abs(ref).max()


def test_volume_is_reference_layout():
    """corr.py:59: [B, h, w, 1, h, w]; entry [b,y1,x1,0,y2,x2] pairs fmap1(y1,x1) with fmap2(y2,x2)."""
    b, d, h, w = 2, 32, 8, 12
    f1 = torch.zeros(b, d, h, w, device=DEV)
    f2 = torch.zeros(b, d, h, w, device=DEV)
    f1[1, :, 3, 5] = 1.0
    f2[1, :, 6, 2] = 2.0
    v = ff().CorrBlock.corr(f1, f2, precision="fp16")
    assert v.shape == (b, h, w, 1, h, w)
    assert abs(float(v[1, 3, 5, 0, 6, 2]) - 2.0 * d / np.sqrt(d)) < 1e-4
    assert int((v != 0).sum()) == 1


# ---------------------------------------------------------------- pyramid
@pytest.mark.parametrize("shape", [(300, 46, 62), (64, 47, 156), (33, 17, 21), (10, 8, 8), (5, 55, 128), (3, 150, 200)])
def test_pyramid_bit_exact_vs_oracle(shape):
    from focusflow_official_b200 import _lib

    q, h, w = shape
    rng = np.random.default_rng(7)
    l0 = (rng.standard_normal(shape) * 19).astype(np.float32)
    nl = 4
    lv = [t(l0[:, None])] + [torch.empty((q, 1, h >> i, w >> i), device=DEV) for i in range(1, nl)]
    _lib.check(_lib.lib().ffcorr_pyramid_f32(_lib.ptr_array(lv), nl, q, h, w, _lib.current_stream()), "pyramid")
    ref = co.pyramid(l0, nl)
    for i in range(1, nl):
        assert np.array_equal(lv[i][:, 0].cpu().numpy(), ref[i]), (shape, i)


def test_pyramid_unaligned_base_and_many_levels():
    from focusflow_official_b200 import _lib

    q, h, w = 37, 33, 35
    rng = np.random.default_rng(8)
    l0 = rng.standard_normal((q, h, w)).astype(np.float32)
    buf = torch.empty(q * h * w + 1, device=DEV)
    buf[1:] = t(l0).flatten()  # 4-byte aligned only -> scalar path
    nl = 6
    lv = [buf[1:].view(q, 1, h, w)] + [torch.empty((q, 1, h >> i, w >> i), device=DEV) for i in range(1, nl)]
    _lib.check(_lib.lib().ffcorr_pyramid_f32(_lib.ptr_array(lv), nl, q, h, w, _lib.current_stream()), "pyramid")
    ref = co.pyramid(l0, nl)
    for i in range(1, nl):
        assert np.array_equal(lv[i][:, 0].cpu().numpy(), ref[i]), i


# ---------------------------------------------------------------- lookup vs oracle
def _coords_cases(rng, b, h, w):
    grid = co.coords_grid(b, h, w)
    n = lambda s: rng.standard_normal(grid.shape).astype(np.float32) * np.float32(s)
    far = grid + n(200.0)
    far[0, :, 0, 0] = [1e7, -1e7]
    far[0, :, 0, 1] = [np.inf, 0]
    return {"grid": grid, "half": grid + np.float32(0.5), "s1": grid + n(1), "s3": grid + n(3), "s20": grid + n(20), "far": far}


@pytest.mark.parametrize("shape,radius,nl", [((1, 46, 62), 4, 4), ((2, 17, 21), 4, 4), ((1, 24, 40), 3, 4),
                                             ((2, 12, 9), 2, 3), ((1, 9, 33), 1, 2), ((3, 8, 8), 4, 1)])
def test_lookup_vs_oracle(shape, radius, nl):
    b, h, w = shape
    rng = np.random.default_rng(99)
    q = b * h * w
    pyr = [rng.standard_normal((q, h >> i, w >> i)).astype(np.float32) for i in range(nl)]
    levels = [t(p[:, None]) for p in pyr]
    for name, c in _coords_cases(rng, b, h, w).items():
        ref = co.lookup(pyr, c, radius)
        got = ff().lookup(levels, t(c), radius).cpu().numpy()
        assert got.shape == ref.shape == (b, nl * (2 * radius + 1) ** 2, h, w)
        assert np.isfinite(got).all(), name
        assert max_rel(got, ref) <= 1e-5, (shape, name, max_rel(got, ref))


@pytest.mark.parametrize("shape", [(2, 46, 62), (1, 47, 156), (2, 17, 21)])
def test_lookup_vs_the_reference_formula_on_torch_cuda(shape):
    """What the reference runs on a GPU: corr.py:29-50 + utils.py:57-71 through ATen's CUDA kernels, whose division by
    the scalar (W-1) is a multiplication by its fp32 reciprocal.  With sampler="cuda" (the default) the lookups (and
    the adjoint) follow that form and agree to 1e-5; sampler="cpu" differs by up to ~3e-5 of max|value| -- the
    same amount by which the reference's own CPU and GPU runs differ.  Both output layouts."""
    m = ff()
    b, h, w = shape
    torch.manual_seed(41)
    n = h * w
    pyr = [torch.randn(b * n, 1, h >> i, w >> i, device=DEV) for i in range(4)]
    dd = torch.linspace(-4, 4, 9, device=DEV)
    delta = torch.stack(torch.meshgrid(dd, dd, indexing="ij"), dim=-1).view(1, 9, 9, 2)
    for sigma in (0.7, 3.0, 25.0):
        coords = m.coords_grid(b, h, w, DEV) + 0.37 + torch.randn(b, 2, h, w, device=DEV) * sigma
        c = coords.permute(0, 2, 3, 1).reshape(b * n, 1, 1, 2)
        outs = []
        for i, lv in enumerate(pyr):
            cl = c / 2 ** i + delta
            hh, ww = lv.shape[-2:]
            xg, yg = cl.split([1, 1], dim=-1)
            grid = torch.cat([2 * xg / (ww - 1) - 1, 2 * yg / (hh - 1) - 1], dim=-1)
            outs.append(torch.nn.functional.grid_sample(lv, grid, align_corners=True).view(b, h, w, -1))
        ref = torch.cat(outs, dim=-1).permute(0, 3, 1, 2).contiguous().float()
        scale = float(ref.abs().max())
        tl = m.tile_levels(pyr)
        for cl_out in (False, True):
            for got in (m.lookup(pyr, coords, 4, sampler="cpu", channels_last=cl_out),
                        m.lookup_tiled(tl, coords, 4, sampler="cpu", channels_last=cl_out)):
                assert float((got - ref).abs().max()) <= 1e-4 * scale, ("cpu", sigma)
            for got in (m.lookup(pyr, coords, 4, sampler="cuda", channels_last=cl_out),
                        m.lookup_tiled(tl, coords, 4, sampler="cuda", channels_last=cl_out)):
                assert float((got - ref).abs().max()) <= 1e-5 * scale, ("cuda", sigma, float((got - ref).abs().max()) / scale)


def test_corrblock_surface_and_errors():
    m = ff()
    f = torch.randn(1, 16, 16, 24, device=DEV)
    blk = m.CorrBlock(f, f)
    assert blk.num_levels == 4 and blk.radius == 4 and len(blk.corr_pyramid) == 4
    assert [tuple(l.shape) for l in blk.corr_pyramid] == [(384, 1, 16, 24), (384, 1, 8, 12), (384, 1, 4, 6), (384, 1, 2, 3)]
    out = blk(m.coords_grid(1, 16, 24, DEV))
    assert out.shape == (1, 324, 16, 24) and out.dtype == torch.float32 and out.is_contiguous()
    with pytest.raises(ValueError):
        blk(torch.zeros(1, 2, 8, 8, device=DEV))
    with pytest.raises(NotImplementedError):
        m.CorrBlock(f.cpu(), f.cpu())
    with pytest.raises(ValueError):
        m.CorrBlock(torch.randn(1, 8, 4, 4, device=DEV), torch.randn(1, 8, 4, 4, device=DEV))  # too small for 4 levels
    with pytest.raises(RuntimeError):
        m.lookup(blk.corr_pyramid, m.coords_grid(1, 16, 24, DEV), radius=7)


def test_empty_batch():
    m = ff()
    f = torch.zeros(0, 32, 16, 16, device=DEV)
    blk = m.CorrBlock(f, f)
    assert blk(torch.zeros(0, 2, 16, 16, device=DEV)).shape == (0, 324, 16, 16)
    assert m.FunctionCorrelation(torch.zeros(0, 8, 5, 5, device=DEV), torch.zeros(0, 8, 5, 5, device=DEV)).shape == (0, 81, 5, 5)


# ---------------------------------------------------------------- full-size properties (BASELINE config 2)
def test_full_size_config2_properties():
    """B=8, 376x1248 -> fmap [8,256,47,156], N=7332 (tile tails in both GEMM dims)."""
    m = ff()
    torch.manual_seed(1234)
    b, d, h, w = 8, 256, 47, 156
    f1 = torch.randn(b, d, h, w, device=DEV) * 4.4
    f2 = torch.randn(b, d, h, w, device=DEV) * 4.4
    blk = m.CorrBlock(f1, f2)
    n = h * w
    l0 = blk.corr_pyramid[0].view(b, n, n)
    # (1) sampled entries against an fp64 dot product
    idx = torch.randint(0, n, (4096, 2), device=DEV)
    bi = torch.randint(0, b, (4096,), device=DEV)
    a = f1.view(b, d, n)[bi, :, idx[:, 0]].double()
    c = f2.view(b, d, n)[bi, :, idx[:, 1]].double()
    ref = (a * c).sum(1) / 16.0
    got = l0[bi, idx[:, 0], idx[:, 1]].double()
    assert float((got - ref).norm() / ref.norm()) <= 1e-3
    # the last row / column of the last batch item (tails of the 128x256 tiling)
    ref_row = (f1.view(b, d, n)[b - 1, :, n - 1].double()[:, None] * f2.view(b, d, n)[b - 1].double()).sum(0) / 16.0
    assert float((l0[b - 1, n - 1].double() - ref_row).norm() / ref_row.norm()) <= 1e-3
    # (2) checksum: sum_j C[i,j] = f1[:,i] . sum_j f2[:,j] / sqrt(D)   (linearity of the contraction)
    colsum = f2.view(b, d, n).double().sum(2)
    ref_sum = (f1.view(b, d, n).double() * colsum[:, :, None]).sum(1) / 16.0
    got_sum = l0.double().sum(2)
    assert float((got_sum - ref_sum).abs().max() / ref_sum.abs().max()) <= 2e-3
    # (3) swapping the operands transposes the volume
    blk_t = m.CorrBlock(f2[:1], f1[:1])
    assert torch.allclose(blk_t.corr_pyramid[0].view(n, n), l0[0].t().contiguous(), rtol=1e-6, atol=1e-5)
    del blk_t
    # (4) pyramid == ATen avg_pool2d bit for bit at full size
    cur = blk.corr_pyramid[0]
    for i in range(1, 4):
        cur = torch.nn.functional.avg_pool2d(cur, 2, stride=2)
        assert torch.equal(cur, blk.corr_pyramid[i]), i
    # (5) lookup at exact integer coordinates reads the volume itself
    coords = m.coords_grid(b, h, w, DEV)
    out = blk(coords)
    assert out.shape == (b, 324, h, w)
    l0m = blk.corr_pyramid[0].view(b, h, w, h, w)
    for (aa, bb) in [(4, 4), (0, 0), (8, 3), (2, 8)]:
        dx, dy = aa - 4, bb - 4
        ys = torch.arange(h, device=DEV)[:, None].expand(h, w)
        xs = torch.arange(w, device=DEV)[None, :].expand(h, w)
        ty, tx = ys + dy, xs + dx
        ok = (ty >= 0) & (ty < h) & (tx >= 0) & (tx < w)
        gathered = l0m[:, ys, xs, ty.clamp(0, h - 1), tx.clamp(0, w - 1)] * ok
        ch = out[:, aa * 9 + bb]
        assert float((ch - gathered).abs().max()) <= 2e-3 * float(l0m.abs().max())
    # (6) idempotence / determinism
    assert torch.equal(out, blk(coords))
    # (7) a subset of queries against the CPU oracle on the real pyramid
    coords_r = coords + torch.randn_like(coords) * 3
    out_r = blk(coords_r)
    sel = torch.arange(0, n, 97, device=DEV)
    pyr = [lv.view(b, n, lv.shape[2], lv.shape[3])[b - 1, sel].cpu().numpy() for lv in blk.corr_pyramid]
    cxy = coords_r.view(b, 2, n)[b - 1][:, sel].cpu().numpy()
    ref_r = np.concatenate([co.lookup_level(pyr[i], cxy[0] / np.float32(2 ** i), cxy[1] / np.float32(2 ** i), 4)
                            for i in range(4)], axis=1)
    got_r = out_r.view(b, 324, n)[b - 1][:, sel].t().cpu().numpy()
    assert max_rel(got_r, ref_r) <= 1e-5


# ---------------------------------------------------------------- AlternateCorrBlock
@pytest.mark.parametrize("shape,nl,radius,chunk", [((2, 64, 24, 40), 4, 4, 128), ((1, 32, 17, 21), 4, 4, 100), ((2, 48, 16, 16), 3, 3, 37),
                                                   ((1, 256, 46, 62), 4, 4, None), ((1, 16, 9, 13), 1, 2, 50), ((1, 32, 33, 47), 4, 4, 1551)])
def test_alternate_corrblock_equals_corrblock_bitwise(shape, nl, radius, chunk):
    """Memory-bounded, recomputed-per-lookup block (chunks of queries) == the stored-pyramid block, bit for bit,
    for chunk sizes that are / are not multiples of the GEMM row tile (the class rounds to whole 32-query lookup units)."""
    m = ff()
    torch.manual_seed(12)
    b, d, h, w = shape
    f1 = torch.randn(shape, device=DEV) * 4.4
    f2 = torch.randn(shape, device=DEV) * 4.4
    ref = m.CorrBlock(f1, f2, num_levels=nl, radius=radius)
    alt = m.AlternateCorrBlock(f1, f2, num_levels=nl, radius=radius, chunk=chunk)
    assert alt.chunk % 32 == 0
    for sigma in (0.0, 2.5, 30.0):
        coords = m.coords_grid(b, h, w, DEV) + torch.randn(b, 2, h, w, device=DEV) * sigma
        a, r = alt(coords), ref(coords)
        assert a.shape == r.shape and torch.isfinite(a).all()
        assert torch.equal(a, r), (sigma, (a != r).sum().item())
    with pytest.raises(ValueError):
        m.AlternateCorrBlock(f1, f2, num_levels=nl, precision="fp32")


@pytest.mark.parametrize("shape,heads", [((2, 64, 16, 24), 1), ((1, 128, 17, 21), 1), ((1, 96, 12, 16), 2)])
def test_flowformer_cost_volume_and_flow_token(shape, heads):
    """FlowFormer's unscaled multi-head volume (encoder.py:335-347) and single-level token lookup (decoder.py:185-203)
    against their literal torch formulas."""
    from focusflow_official_b200 import flowformer as fl

    torch.manual_seed(3)
    b, dim, h, w = shape
    f1 = torch.randn(shape, device=DEV)
    f2 = torch.randn(shape, device=DEV)
    d = dim // heads
    a = f1.view(b, heads, d, h * w).permute(0, 1, 3, 2).double()
    c = f2.view(b, heads, d, h * w).permute(0, 1, 3, 2).double()
    ref = torch.einsum("bhid,bhjd->bhij", a, c).view(b, heads, h, w, h, w)
    for prec, tol in (("fp16", 1e-3), ("fp32", 2e-6)):
        got = fl.cost_volume(f1, f2, heads=heads, precision=prec)
        assert got.shape == ref.shape
        assert float((got.double() - ref).norm() / ref.norm()) <= tol, prec
    # token lookup on the reference layout [B*h*w, heads, h, w]
    vol = fl.cost_volume(f1, f2, heads=heads, precision="fp32")
    cost_maps = vol.permute(0, 2, 3, 1, 4, 5).reshape(b * h * w, heads, h, w).contiguous()
    coords = ff().coords_grid(b, h, w, DEV) + torch.randn(b, 2, h, w, device=DEV) * 2
    got = fl.encode_flow_token(cost_maps, coords)
    dd = torch.linspace(-4, 4, 9, device=DEV)
    delta = torch.stack(torch.meshgrid(dd, dd, indexing="ij"), dim=-1).view(1, 9, 9, 2)
    cl = coords.permute(0, 2, 3, 1).reshape(b * h * w, 1, 1, 2) + delta
    grid = torch.stack([2 * cl[..., 0] / (w - 1) - 1, 2 * cl[..., 1] / (h - 1) - 1], -1)
    want = torch.nn.functional.grid_sample(cost_maps, grid, align_corners=True).view(b, h, w, -1).permute(0, 3, 1, 2)
    assert got.shape == want.shape == (b, heads * 81, h, w)
    assert float((got - want).abs().max()) <= 2e-5 * float(want.abs().max())


# ---------------------------------------------------------------- gradients
def test_corrblock_gradients_match_torch_autograd():
    m = ff()
    torch.manual_seed(5)
    b, d, h, w = 2, 32, 16, 20
    f1 = torch.randn(b, d, h, w, device=DEV, requires_grad=True)
    f2 = torch.randn(b, d, h, w, device=DEV, requires_grad=True)
    coords = m.coords_grid(b, h, w, DEV) + torch.randn(b, 2, h, w, device=DEV) * 2
    wgt = torch.randn(b, 324, h, w, device=DEV)

    blk = m.CorrBlock(f1, f2, precision="fp32")
    out = blk(coords.detach())
    (out * wgt).sum().backward()
    g1, g2 = f1.grad.clone(), f2.grad.clone()
    f1.grad = None
    f2.grad = None

    # pure-torch restatement of corr.py (same formulas as the reference) as the autograd checker
    n = h * w
    corr = torch.matmul(f1.view(b, d, n).transpose(1, 2), f2.view(b, d, n)) / torch.sqrt(torch.tensor(float(d)))
    cur = corr.reshape(b * n, 1, h, w)
    pyr = [cur]
    for _ in range(3):
        cur = torch.nn.functional.avg_pool2d(cur, 2, stride=2)
        pyr.append(cur)
    c = coords.permute(0, 2, 3, 1).reshape(b * n, 1, 1, 2)
    dd = torch.linspace(-4, 4, 9, device=DEV)
    delta = torch.stack(torch.meshgrid(dd, dd, indexing="ij"), dim=-1).view(1, 9, 9, 2)
    outs = []
    for i, lv in enumerate(pyr):
        cl = c / 2 ** i + delta
        hh, ww = lv.shape[-2:]
        gx = 2 * cl[..., 0] / (ww - 1) - 1
        gy = 2 * cl[..., 1] / (hh - 1) - 1
        s = torch.nn.functional.grid_sample(lv, torch.stack([gx, gy], -1), align_corners=True)
        outs.append(s.view(b, h, w, -1))
    ref = torch.cat(outs, -1).permute(0, 3, 1, 2)
    assert float((ref - out).abs().max()) <= 1e-4 * float(ref.abs().max())
    (ref * wgt).sum().backward()
    assert float((g1 - f1.grad).norm() / f1.grad.norm()) <= 1e-4
    assert float((g2 - f2.grad).norm() / f2.grad.norm()) <= 1e-4


@pytest.mark.parametrize("shape", [(2, 32, 16, 20), (1, 256, 46, 62), (1, 24, 9, 13), (2, 40, 8, 8)])
@pytest.mark.parametrize("precision", ["fp16", "tf32"])
def test_tensor_core_backward_gemms(shape, precision):
    """precision != fp32: the two backward GEMMs run on tcgen05 (tf32 operands) when h*w % 4 == 0; they must agree
    with the exact CUDA-core path to tf32 accuracy, for the volume gradient alone and through a lookup."""
    m = ff()
    torch.manual_seed(9)
    b, d, h, w = shape
    f1 = (torch.randn(shape, device=DEV) * 2).requires_grad_(True)
    f2 = (torch.randn(shape, device=DEV) * 2).requires_grad_(True)
    nl = 4 if min(h, w) >= 8 else 1
    coords = m.coords_grid(b, h, w, DEV) + torch.randn(b, 2, h, w, device=DEV) * 2
    wgt = torch.randn(b, nl * 81, h, w, device=DEV)
    wl0 = torch.randn(b * h * w, 1, h, w, device=DEV)
    grads = {}
    for prec in (precision, "fp32"):
        f1.grad = f2.grad = None
        blk = m.CorrBlock(f1, f2, num_levels=nl, precision=prec)
        ((blk(coords) * wgt).sum() + (blk.corr_pyramid[0] * wl0).sum()).backward()
        grads[prec] = (f1.grad.clone(), f2.grad.clone())
    for got, ref in zip(grads[precision], grads["fp32"]):
        assert torch.isfinite(got).all()
        assert float((got - ref).norm() / ref.norm()) <= 2e-3, float((got - ref).norm() / ref.norm())


def test_autograd_block_with_tiled_forward_equals_rowmajor_block():
    """Under autograd the forward pyramid is stored tiled (fused build) when a tensor-core precision is used; outputs
    and gradients must equal those of the row-major block, including a level used directly in the loss."""
    m = ff()
    torch.manual_seed(17)
    b, d, h, w = 2, 48, 24, 32
    f1 = (torch.randn(b, d, h, w, device=DEV) * 2).requires_grad_(True)
    f2 = (torch.randn(b, d, h, w, device=DEV) * 2).requires_grad_(True)
    coords = [m.coords_grid(b, h, w, DEV) + torch.randn(b, 2, h, w, device=DEV) * 2 for _ in range(4)]
    wgt = [torch.randn(b, 324, h, w, device=DEV) for _ in range(4)]
    wl2 = torch.randn(b * h * w, 1, h // 4, w // 4, device=DEV)
    res = {}
    for layout in ("tiled", "rowmajor"):
        f1.grad = f2.grad = None
        blk = m.CorrBlock(f1, f2, precision="fp16", layout=layout)
        assert blk._grad_tiled == (layout == "tiled")
        outs = [blk(c) for c in coords]
        loss = sum((o * wg).sum() for o, wg in zip(outs, wgt)) + (blk.corr_pyramid[2] * wl2).sum()
        loss.backward()
        res[layout] = ([o.detach() for o in outs], f1.grad.clone(), f2.grad.clone())
    for a, r in zip(res["tiled"][0], res["rowmajor"][0]):
        assert torch.equal(a, r)
    for k in (1, 2):
        a, r = res["tiled"][k], res["rowmajor"][k]
        assert float((a - r).norm() / r.norm()) <= 1e-5, k     # same terms, different accumulation order in the sink


def test_corrblock_gradients_many_lookups_and_direct_level_use():
    """12 lookups share one gradient buffer (_GradSink); a level used directly in the loss adds to it; a second
    backward through the retained graph starts from zero again."""
    m = ff()
    torch.manual_seed(6)
    b, d, h, w = 1, 16, 16, 24
    n = h * w
    f1 = torch.randn(b, d, h, w, device=DEV, requires_grad=True)
    f2 = torch.randn(b, d, h, w, device=DEV, requires_grad=True)
    coords = [m.coords_grid(b, h, w, DEV) + torch.randn(b, 2, h, w, device=DEV) * 2 for _ in range(12)]
    wgt = [torch.randn(b, 324, h, w, device=DEV) for _ in range(12)]
    wl1 = torch.randn(b * n, 1, h // 2, w // 2, device=DEV)

    def ref_loss():
        corr = torch.matmul(f1.view(b, d, n).transpose(1, 2), f2.view(b, d, n)) / torch.sqrt(torch.tensor(float(d)))
        cur = corr.reshape(b * n, 1, h, w)
        pyr = [cur]
        for _ in range(3):
            cur = torch.nn.functional.avg_pool2d(cur, 2, stride=2)
            pyr.append(cur)
        dd = torch.linspace(-4, 4, 9, device=DEV)
        delta = torch.stack(torch.meshgrid(dd, dd, indexing="ij"), dim=-1).view(1, 9, 9, 2)
        loss = (pyr[1] * wl1).sum()
        for cds, wg in zip(coords, wgt):
            c = cds.permute(0, 2, 3, 1).reshape(b * n, 1, 1, 2)
            outs = []
            for i, lv in enumerate(pyr):
                cl = c / 2 ** i + delta
                hh, ww = lv.shape[-2:]
                grid = torch.stack([2 * cl[..., 0] / (ww - 1) - 1, 2 * cl[..., 1] / (hh - 1) - 1], -1)
                outs.append(torch.nn.functional.grid_sample(lv, grid, align_corners=True).view(b, h, w, -1))
            loss = loss + (torch.cat(outs, -1).permute(0, 3, 1, 2) * wg).sum()
        return loss

    blk = m.CorrBlock(f1, f2, precision="fp32")
    loss = (blk.corr_pyramid[1] * wl1).sum()
    for cds, wg in zip(coords, wgt):
        loss = loss + (blk(cds) * wg).sum()
    loss.backward(retain_graph=True)
    g1, g2 = f1.grad.clone(), f2.grad.clone()
    f1.grad = f2.grad = None
    loss.backward()                      # second pass through the same graph: same gradients, not doubled
    assert torch.allclose(f1.grad, g1, rtol=1e-5, atol=1e-5) and torch.allclose(f2.grad, g2, rtol=1e-5, atol=1e-5)
    f1.grad = f2.grad = None
    ref_loss().backward()
    assert float((g1 - f1.grad).norm() / f1.grad.norm()) <= 1e-4
    assert float((g2 - f2.grad).norm() / f2.grad.norm()) <= 1e-4


def test_lookup_backward_vs_oracle_adjoint():
    from focusflow_official_b200 import _lib

    rng = np.random.default_rng(21)
    b, h, w = 1, 10, 12
    q = b * h * w
    coords = co.coords_grid(b, h, w) + rng.standard_normal((b, 2, h, w)).astype(np.float32) * 3
    gout = rng.standard_normal((b, 2 * 81, h, w)).astype(np.float32)
    glv = [torch.zeros(q, 1, h >> i, w >> i, device=DEV) for i in range(2)]
    c_dev, g_dev = t(coords), t(gout)  # keep alive: data_ptr() of a temporary dangles
    _lib.check(_lib.lib().ffcorr_lookup_bwd_f32(_lib.ptr_array(glv), 2, c_dev.data_ptr(), g_dev.data_ptr(), b, h, w, 4,
                                                0, _lib.current_stream()), "bwd")
    torch.cuda.synchronize()
    c = coords.transpose(0, 2, 3, 1).reshape(q, 2)
    g = gout.transpose(0, 2, 3, 1).reshape(q, 2, 81)
    for i in range(2):
        ref = co.lookup_level_backward(g[:, i], (q, h >> i, w >> i), c[:, 0] / np.float32(2 ** i), c[:, 1] / np.float32(2 ** i), 4)
        got = glv[i][:, 0].cpu().numpy()
        assert np.abs(got - ref).max() <= 1e-4 * max(1.0, np.abs(ref).max()), i


@pytest.mark.parametrize("shape,radius,nl", [((1, 46, 62), 4, 4), ((2, 17, 21), 4, 4), ((2, 12, 9), 2, 3), ((1, 9, 33), 1, 2),
                                             ((1, 24, 40), 3, 4)])
def test_lookup_backward_is_the_adjoint_of_the_forward(shape, radius, nl):
    """<lookup(P, c), G> == <P, lookup_bwd(G, c)> for integer (exact per-tap path), half-integer, noisy and far
    out-of-bounds / non-finite coordinates; accumulation over two calls doubles the gradient."""
    from focusflow_official_b200 import _lib

    m = ff()
    b, h, w = shape
    rng = np.random.default_rng(31)
    q = b * h * w
    k2 = (2 * radius + 1) ** 2
    pyr = [t(rng.standard_normal((q, 1, h >> i, w >> i)).astype(np.float32)) for i in range(nl)]
    for name, c in _coords_cases(rng, b, h, w).items():
        c_dev = t(c)
        gout = t(rng.standard_normal((b, nl * k2, h, w)).astype(np.float32))
        out = m.lookup(pyr, c_dev, radius)
        lhs = float((out.double() * gout.double()).sum())
        glv = [torch.zeros_like(p) for p in pyr]
        for _ in range(2):
            _lib.check(_lib.lib().ffcorr_lookup_bwd_f32(_lib.ptr_array(glv), nl, c_dev.data_ptr(), gout.data_ptr(), b, h, w,
                                                        radius, 1, _lib.current_stream()), "bwd")
        torch.cuda.synchronize()
        rhs = sum(float((p.double() * g.double()).sum()) for p, g in zip(pyr, glv)) / 2
        scale = float(out.double().abs().mul(gout.double().abs()).sum()) + 1e-30
        assert all(torch.isfinite(g).all() for g in glv), name
        assert abs(lhs - rhs) <= 1e-5 * scale, (name, lhs, rhs)


# ---------------------------------------------------------------- PWC cost volume
PWC_SHAPES = [(2, 32, 28, 64), (1, 64, 56, 128), (2, 96, 28, 64), (2, 128, 14, 32), (2, 196, 7, 16),
              (1, 3, 5, 5), (2, 40, 9, 13), (1, 17, 33, 70)]


@pytest.mark.parametrize("shape", PWC_SHAPES)
def test_pwc_forward_vs_oracle(shape):
    rng = np.random.default_rng(3)
    one = rng.standard_normal(shape).astype(np.float32)
    two = rng.standard_normal(shape).astype(np.float32)
    ref = po.forward_c(one, two)
    got = ff().FunctionCorrelation(tenOne=t(one), tenTwo=t(two)).cpu().numpy()
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max()), shape
    fused = ff().correlation_leaky(t(one), t(two), 0.1).cpu().numpy()
    assert np.array_equal(fused, np.where(got > 0, got, got * np.float32(0.1)).astype(np.float32))


def test_pwc_full_size_level2_properties():
    """Config 3, pyramid level 2: B=16, C=32, 112x256."""
    m = ff()
    torch.manual_seed(3)
    one = torch.randn(16, 32, 112, 256, device=DEV)
    two = torch.randn(16, 32, 112, 256, device=DEV)
    out = m.FunctionCorrelation(one, two)
    assert out.shape == (16, 81, 112, 256)
    # centre channel is the plain channel mean of one*two; shifted channels are shifted products
    assert float((out[:, 40] - (one * two).mean(1)).abs().max()) <= 1e-5
    dy, dx = 3, -2
    ref = (one[:, :, : 112 - dy, -dx:] * two[:, :, dy:, : 256 + dx]).mean(1)
    assert float((out[:, (dy + 4) * 9 + dx + 4, : 112 - dy, -dx:] - ref).abs().max()) <= 1e-5
    assert float(out[:, (dy + 4) * 9 + dx + 4, 112 - dy:, :].abs().max()) == 0.0
    # linearity in the second argument
    out2 = m.FunctionCorrelation(one, two * 2.0)
    assert float((out2 - 2 * out).abs().max()) <= 1e-5
    # swap symmetry: corr(one,two)[dy,dx](y,x) == corr(two,one)[-dy,-dx](y+dy,x+dx)
    outs = m.FunctionCorrelation(two, one)
    lhs = out[:, (dy + 4) * 9 + dx + 4, : 112 - dy, -dx:]
    rhs = outs[:, (-dy + 4) * 9 + (-dx) + 4, dy:, : 256 + dx]
    assert float((lhs - rhs).abs().max()) <= 1e-5


@pytest.mark.parametrize("shape", [(2, 20, 11, 14),      # W % 4 != 0: per-element fallback kernel
                                   (2, 20, 11, 16),      # TMA kernel, one partial channel stage, image < tile
                                   (1, 70, 19, 40),      # three channel stages (32+32+6), several tiles, ragged rows
                                   (2, 32, 8, 32),       # exactly one tile, power-of-two C (reciprocal multiply)
                                   (1, 196, 7, 16)])     # config-3 level 6
def test_pwc_backward_vs_oracle(shape):
    m = ff()
    rng = np.random.default_rng(4)
    one = rng.standard_normal(shape).astype(np.float32)
    two = rng.standard_normal(shape).astype(np.float32)
    g = rng.standard_normal((shape[0], 81, shape[2], shape[3])).astype(np.float32)
    a = t(one).requires_grad_(True)
    bb = t(two).requires_grad_(True)
    m.ModuleCorrelation()(a, bb).backward(t(g))
    r1, r2 = po.backward_c(one, two, g)
    assert np.abs(a.grad.cpu().numpy() - r1).max() <= 1e-5 * max(1.0, np.abs(r1).max())
    assert np.abs(bb.grad.cpu().numpy() - r2).max() <= 1e-5 * max(1.0, np.abs(r2).max())
    with pytest.raises(NotImplementedError):
        m.FunctionCorrelation(torch.zeros(1, 2, 4, 4), torch.zeros(1, 2, 4, 4))


# ---------------------------------------------------------------- end-to-end EPE (BASELINE: <= 0.01 px after 12 iters)
def test_e2e_epe_vs_reference_golden():
    """The PyTorch host model + B200 CorrBlock against FF_RAFT_FUSION outputs recorded by
    oracle/make_golden.py (reference on CPU, same deterministic weights and inputs)."""
    import sys

    sys.path.insert(0, os.path.dirname(__file__))
    from test_host_model import make_model
    from weights import synthetic_pair

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ffraft_e2e.npz"))
    torch.backends.cudnn.allow_tf32 = False  # isolate the correlation path: fp32 convs like the CPU reference
    torch.backends.cuda.matmul.allow_tf32 = False
    model = make_model(DEV)
    for tag in ("a", "b"):
        b, hh, ww, iters = [int(v) for v in g[f"{tag}_shape"]]
        im1, im2, m1, m2 = (x.to(DEV) for x in synthetic_pair(b, hh, ww, seed=1234 + b))
        for prec in ("fp16", "fp32"):
            model.flow_net.corr_precision = prec
            with torch.no_grad():
                lo, up = model(im1, im2, m1, m2, raft_iters=iters, test_mode=True)
            epe = torch.linalg.norm(up.cpu() - torch.from_numpy(g[f"{tag}_flow_up"]), dim=1)
            print(f"e2e {tag} {prec}: EPE mean {float(epe.mean()):.2e} max {float(epe.max()):.2e}")
            assert float(epe.max()) <= 1e-2, (tag, prec, float(epe.max()))  # BASELINE: EPE within 0.01 px
            assert float(epe.mean()) <= 2e-3


# ---------------------------------------------------------------- tiled ("T4") layout: the inference fast path
@pytest.mark.parametrize("shape", [(3, 16, 24), (2, 17, 21), (1, 46, 62), (2, 9, 33), (1, 5, 7)])
def test_tile_untile_roundtrip_and_padding(shape):
    m = ff()
    q, h, w = shape
    x = torch.randn(q, 1, h, w, device=DEV)
    (tl,) = m.tile_levels([x])
    th, tw = (h + 3) // 4, ((w + 3) // 4 + 1) & ~1      # tile-row pitch is even (128-byte aligned tile rows)
    assert tl.shape == (q, th * tw * 16)
    # padding is exact zero and the sum is preserved
    assert float(tl.double().sum()) == pytest.approx(float(x.double().sum()), rel=1e-9, abs=1e-9)
    from focusflow_official_b200 import _lib
    back = torch.empty_like(x)
    _lib.check(_lib.lib().ffcorr_untile_f32(tl.data_ptr(), back.data_ptr(), q, h, w, _lib.current_stream()), "untile")
    assert torch.equal(back, x)
    # explicit element check of the tile-major order
    tt = tl.view(q, th, tw, 4, 4)
    y, xx = min(h - 1, 6), min(w - 1, 5)
    assert torch.equal(tt[:, y // 4, xx // 4, y % 4, xx % 4], x[:, 0, y, xx])


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_golden_lookup_tiled(path):
    """Golden pyramids of the reference, re-laid out in tiles, through the tiled lookup kernel."""
    g = np.load(path)
    m = ff()
    levels = m.tile_levels([t(g[f"level{i}"][:, None]) for i in range(4)])
    for k in [k[len("coords_"):] for k in g.files if k.startswith("coords_")]:
        out = m.lookup_tiled(levels, t(g[f"coords_{k}"]), 4).cpu().numpy()
        ref = g[f"lookup_{k}"]
        assert max_rel(out, ref) <= 1e-5, (k, max_rel(out, ref))


@pytest.mark.parametrize("shape,radius,nl", [((1, 46, 62), 4, 4), ((2, 17, 21), 4, 4), ((1, 24, 40), 3, 4),
                                             ((2, 12, 9), 2, 3), ((1, 9, 33), 1, 2), ((3, 8, 8), 4, 1)])
def test_lookup_tiled_vs_oracle(shape, radius, nl):
    b, h, w = shape
    rng = np.random.default_rng(77)
    q = b * h * w
    pyr = [rng.standard_normal((q, h >> i, w >> i)).astype(np.float32) for i in range(nl)]
    levels = ff().tile_levels([t(p[:, None]) for p in pyr])
    for name, c in _coords_cases(rng, b, h, w).items():
        ref = co.lookup(pyr, c, radius)
        got = ff().lookup_tiled(levels, t(c), radius).cpu().numpy()
        assert np.isfinite(got).all(), name
        assert max_rel(got, ref) <= 1e-5, (shape, name, max_rel(got, ref))


@pytest.mark.parametrize("precision", ["fp16", "tf32", "bf16x3"])
@pytest.mark.parametrize("shape,nl", [((1, 64, 47, 156), 4), ((2, 32, 33, 47), 4), ((1, 256, 46, 62), 4), ((2, 40, 17, 21), 4),
                                      ((1, 32, 16, 16), 4), ((1, 32, 8, 8), 4), ((3, 16, 24, 40), 3), ((1, 48, 9, 35), 2),
                                      ((1, 16, 50, 20), 4), ((2, 24, 12, 9), 1)])
def test_fused_build_equals_volume_then_pyramid_bitwise(shape, nl, precision):
    """ffcorr_build_tiled_f32 (pooling in the GEMM epilogue) == ffcorr_volume_tiled_f32 + ffcorr_pyramid_tiled_f32,
    bit for bit, INCLUDING the zero padding of every tile (raw tiled buffers are compared)."""
    m = ff()
    torch.manual_seed(5)
    f1 = torch.randn(shape, device=DEV) * 4.4
    f2 = torch.randn(shape, device=DEV) * 4.4
    b, d, h, w = shape
    # poison the allocator's blocks so stale data would show up in unwritten padding
    junk = [torch.full((b * h * w, int(m._lib.lib().ffcorr_tiled_map_elems(h, w, i))), float("nan"), device=DEV) for i in range(nl)]
    del junk
    fused = m.tiled_pyramid(f1, f2, nl, precision, fused=True)
    junk = [torch.full_like(x, float("nan")) for x in fused]
    del junk
    plain = m.tiled_pyramid(f1, f2, nl, precision, fused=False)
    torch.cuda.synchronize()
    for i in range(nl):
        hi, wi = h >> i, w >> i
        th, twr = (hi + 3) // 4, (wi + 3) // 4
        # the extra tile column that keeps the pitch even is never read (content unspecified); compare the rest raw
        fa = fused[i].view(b * h * w, th, -1, 16)[:, :, :twr]
        pa = plain[i].view(b * h * w, th, -1, 16)[:, :, :twr]
        assert torch.isfinite(fa).all(), i
        assert torch.equal(fa, pa), (i, (fa != pa).sum().item())


@pytest.mark.parametrize("shape", [(1, 256, 46, 62), (2, 64, 17, 21), (2, 32, 16, 24), (1, 40, 9, 13),
                                   (1, 32, 100, 160)])   # 64 KB maps: beyond the standalone tiled pyramid, fused build only
def test_tiled_and_rowmajor_blocks_agree_bitwise(shape):
    """Same kernels' arithmetic, two storage orders: pyramids and lookups must be identical."""
    m = ff()
    b, d, h, w = shape
    nl = 4 if min(h, w) >= 16 else 1
    torch.manual_seed(2)
    f1 = torch.randn(shape, device=DEV) * 4.4
    f2 = torch.randn(shape, device=DEV) * 4.4
    a = m.CorrBlock(f1, f2, num_levels=nl, layout="tiled")
    r = m.CorrBlock(f1, f2, num_levels=nl, layout="rowmajor")
    assert a._tiled and not r._tiled
    for i in range(nl):
        assert torch.equal(a.corr_pyramid[i], r.corr_pyramid[i]), i
    coords = m.coords_grid(b, h, w, DEV) + torch.randn(b, 2, h, w, device=DEV) * 3
    assert torch.equal(a(coords), r(coords))
    grid = m.coords_grid(b, h, w, DEV)
    assert torch.equal(a(grid), r(grid))


# ---------------------------------------------------------------- randomized shape sweep
def test_random_shapes_tiled_rowmajor_alternate_agree():
    """40 random (B, D, h, w, levels, radius): fused tiled block == row-major block == chunked alternate block bit for
    bit, level 0 within the fp16-operand bar of an fp64 contraction, lookups within 1e-5 of the numpy oracle on the
    block's own pyramid (spot-checked on a subset of queries)."""
    m = ff()
    rng = np.random.default_rng(2024)
    for it in range(40):
        h, w = int(rng.integers(8, 72)), int(rng.integers(8, 72))
        nl = int(rng.integers(1, 5))
        while (h >> (nl - 1)) < 1 or (w >> (nl - 1)) < 1:
            nl -= 1
        b = int(rng.integers(1, 4))
        d = int(rng.choice([8, 24, 64, 100, 256]))
        radius = int(rng.integers(1, 5))
        torch.manual_seed(it)
        f1 = torch.randn(b, d, h, w, device=DEV) * 3
        f2 = torch.randn(b, d, h, w, device=DEV) * 3
        coords = m.coords_grid(b, h, w, DEV) + torch.randn(b, 2, h, w, device=DEV) * float(rng.choice([0.0, 1.5, 12.0]))
        tag = (it, b, d, h, w, nl, radius)
        tl = m.CorrBlock(f1, f2, num_levels=nl, radius=radius, layout="tiled")
        rm = m.CorrBlock(f1, f2, num_levels=nl, radius=radius, layout="rowmajor")
        assert tl._tiled and not rm._tiled, tag
        out_t, out_r = tl(coords), rm(coords)
        assert torch.equal(out_t, out_r), tag
        for i in range(nl):
            assert torch.equal(tl.corr_pyramid[i], rm.corr_pyramid[i]), (tag, i)
        alt = m.AlternateCorrBlock(f1, f2, num_levels=nl, radius=radius, chunk=int(rng.choice([32, 96, 128, 1000])))
        assert torch.equal(alt(coords), out_t), tag
        n = h * w
        ref0 = torch.matmul(f1.view(b, d, n).transpose(1, 2).double(), f2.view(b, d, n).double()) / np.sqrt(d)
        got0 = rm.corr_pyramid[0].view(b, n, n).double()
        assert float((got0 - ref0).norm() / ref0.norm()) <= 1e-3, tag
        if it % 4 == 0:
            sel = np.arange(0, n, max(1, n // 24))
            pyr = [lv.view(b, n, lv.shape[2], lv.shape[3])[0, sel].cpu().numpy() for lv in rm.corr_pyramid]
            cxy = coords.view(b, 2, n)[0][:, sel].cpu().numpy()
            ref = np.concatenate([co.lookup_level(pyr[i], cxy[0] / np.float32(2 ** i), cxy[1] / np.float32(2 ** i), radius)
                                  for i in range(nl)], axis=1)
            got = out_r.view(b, -1, n)[0][:, sel].t().cpu().numpy()
            assert max_rel(got, ref) <= 1e-5, tag


def _pwc_seeded(idx, shape):
    rng = np.random.default_rng(1000 + idx)       # as oracle/make_golden_pwc.py
    b, c, h, w = shape
    return (rng.standard_normal(shape).astype(np.float32), rng.standard_normal(shape).astype(np.float32),
            rng.standard_normal((b, 81, h, w)).astype(np.float32))


def test_pwc_against_reference_cuda_kernel_golden():
    """Forward and both gradients against tests/golden/pwc_ref_cuda.npz = outputs of the reference's own CUDA kernels."""
    m = ff()
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "pwc_ref_cuda.npz"))
    for idx in sorted(int(k.split("_")[1]) for k in g.files if k.startswith("shape_")):
        shape = tuple(int(v) for v in g[f"shape_{idx}"])
        one, two, gout = _pwc_seeded(idx, shape)
        a, bb = t(one).requires_grad_(True), t(two).requires_grad_(True)
        out = m.FunctionCorrelation(a, bb)
        out.backward(t(gout))
        for name, got, ref in (("out", out.detach(), g[f"out_{idx}"]), ("grad_one", a.grad, g[f"gone_{idx}"]),
                               ("grad_two", bb.grad, g[f"gtwo_{idx}"])):
            err = np.abs(got.cpu().numpy() - ref).max()
            assert err <= 1e-5 * max(1.0, np.abs(ref).max()), (idx, name, err)


def test_pwc_against_reference_cuda_kernels_live():
    """The reference's CUDA kernels (oracle/_ref/libpwc_ref_cuda.so, built from correlation.py by
    oracle/build_pwc_ref_cuda.py) run next to ours on this GPU, fresh random inputs, all five built shapes."""
    from oracle import pwc_ref_cuda as ref

    if not ref.available():
        pytest.skip("oracle/_ref/libpwc_ref_cuda.so not built (needs /root/reference at build time)")
    m = ff()
    for idx, shape in enumerate(ref.shapes()):
        torch.manual_seed(500 + idx)
        b, c, h, w = shape
        one = torch.randn(shape, device=DEV)
        two = torch.randn(shape, device=DEV)
        gout = torch.randn(b, 81, h, w, device=DEV)
        r_out, r_g1, r_g2 = ref.run(idx, one, two, gout)
        a, bb = one.clone().requires_grad_(True), two.clone().requires_grad_(True)
        out = m.FunctionCorrelation(a, bb)
        out.backward(gout)
        for name, got, want in (("out", out.detach(), r_out), ("grad_one", a.grad, r_g1), ("grad_two", bb.grad, r_g2)):
            err = float((got - want).abs().max())
            assert err <= 1e-5 * max(1.0, float(want.abs().max())), (shape, name, err)


def test_random_shapes_pwc_forward_backward():
    """Random PWC shapes (aligned and W % 4 != 0, C around the 8 / 32 channel stage sizes) against the C restatement."""
    m = ff()
    rng = np.random.default_rng(77)
    for it in range(12):
        b = int(rng.integers(1, 3))
        c = int(rng.choice([3, 8, 31, 32, 33, 70]))
        h = int(rng.integers(3, 30))
        w = int(rng.integers(3, 50)) if it % 3 else int(rng.integers(1, 12)) * 4
        one = rng.standard_normal((b, c, h, w)).astype(np.float32)
        two = rng.standard_normal((b, c, h, w)).astype(np.float32)
        g = rng.standard_normal((b, 81, h, w)).astype(np.float32)
        a, bb = t(one).requires_grad_(True), t(two).requires_grad_(True)
        out = m.FunctionCorrelation(a, bb)
        out.backward(t(g))
        ref = po.forward_c(one, two)
        r1, r2 = po.backward_c(one, two, g)
        tag = (it, b, c, h, w)
        assert np.abs(out.detach().cpu().numpy() - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max()), tag
        assert np.abs(a.grad.cpu().numpy() - r1).max() <= 1e-5 * max(1.0, np.abs(r1).max()), tag
        assert np.abs(bb.grad.cpu().numpy() - r2).max() <= 1e-5 * max(1.0, np.abs(r2).max()), tag


def test_kernels_run_on_the_current_stream():
    """Everything is enqueued on torch's CURRENT stream (the reference's CuPy kernels ignore it): producing the
    inputs on a side stream and consuming the results there must be correct without any extra synchronisation."""
    m = ff()
    torch.manual_seed(8)
    b, d, h, w = 2, 64, 24, 32
    side = torch.cuda.Stream()
    torch.cuda.synchronize()
    with torch.cuda.stream(side):
        f1 = torch.randn(b, d, h, w, device=DEV) * 3
        f2 = torch.randn(b, d, h, w, device=DEV) * 3
        for _ in range(20):                       # keep the side stream busy so a default-stream launch would race ahead
            f1 = f1 * 1.0001
            f2 = f2 * 0.9999
        coords = m.coords_grid(b, h, w, DEV) + torch.randn(b, 2, h, w, device=DEV) * 2
        blk = m.CorrBlock(f1, f2)
        out_side = blk(coords)
        cv_side = m.FunctionCorrelation(f1[:, :32], f2[:, :32])
    side.synchronize()
    ref = m.CorrBlock(f1, f2)
    assert torch.equal(out_side, ref(coords))
    assert torch.equal(cv_side, m.FunctionCorrelation(f1[:, :32], f2[:, :32]))


def test_config4_on_one_gpu_batch64():
    """SURVEY config 4 on ONE GPU: B = 64 Sintel-shaped pairs, a 16.7 GB pyramid (3.2 G elements: 64-bit offsets,
    blockIdx.z = 2B in the pre-pass, 14 k lookup blocks).  The last batch item must equal the same pair run alone."""
    m = ff()
    free, _ = torch.cuda.mem_get_info()
    if free < 40 << 30:
        pytest.skip("needs ~25 GB of free HBM")
    torch.manual_seed(64)
    b, d, h, w = 64, 256, 55, 128
    f1 = torch.randn(b, d, h, w, device=DEV) * 4.4
    f2 = torch.randn(b, d, h, w, device=DEV) * 4.4
    coords = m.coords_grid(b, h, w, DEV) + torch.randn(b, 2, h, w, device=DEV) * 3
    blk = m.CorrBlock(f1, f2)
    assert blk._tiled and sum(l.numel() for l in blk._levels) > 3_000_000_000
    out = blk(coords)
    torch.cuda.synchronize()
    for j in (0, 37, 63):
        one = m.CorrBlock(f1[j:j + 1], f2[j:j + 1])
        assert torch.equal(one(coords[j:j + 1]), out[j:j + 1]), j
    del blk, out
    torch.cuda.empty_cache()


def test_more_than_four_levels_uses_the_rowmajor_path():
    """num_levels up to 8 is legal (FFCORR_MAX_LEVELS); beyond 4 the block stores row-major levels (generic one-level
    pooling kernel) and must still match the oracle; gradients flow as well."""
    m = ff()
    torch.manual_seed(23)
    b, d, h, w = 1, 32, 32, 48
    nl = 5
    f1 = (torch.randn(b, d, h, w, device=DEV) * 2).requires_grad_(True)
    f2 = torch.randn(b, d, h, w, device=DEV) * 2
    coords = m.coords_grid(b, h, w, DEV) + torch.randn(b, 2, h, w, device=DEV) * 2
    with torch.no_grad():
        blk = m.CorrBlock(f1, f2, num_levels=nl, radius=3)
        assert not blk._tiled and len(blk.corr_pyramid) == nl
        out = blk(coords)
    pyr = [lv[:, 0].cpu().numpy() for lv in blk.corr_pyramid]
    own = co.pyramid(pyr[0], nl)
    assert all(np.array_equal(own[i], pyr[i]) for i in range(nl))
    ref = co.lookup(pyr, coords.cpu().numpy(), 3)
    assert out.shape == (b, nl * 49, h, w)
    assert max_rel(out.cpu().numpy(), ref) <= 1e-5
    blk_g = m.CorrBlock(f1, f2, num_levels=nl, radius=3)
    blk_g(coords).square().sum().backward()
    assert torch.isfinite(f1.grad).all() and float(f1.grad.abs().sum()) > 0


def test_host_model_with_alternate_corr_gives_the_same_flow():
    """raft.py:195-196 `alternate_corr`: the memory-bounded block behind the same host model -> identical flow."""
    from test_host_model import make_model
    from weights import synthetic_pair

    model = make_model().to(DEV).eval()
    im1, im2, m1, _ = (x.to(DEV) for x in synthetic_pair(1, 128, 192, seed=3))
    with torch.no_grad():
        ref = model(im1, im2, m1, None, raft_iters=4, test_mode=True)[1]
        model.flow_net.alternate_corr = True
        alt = model(im1, im2, m1, None, raft_iters=4, test_mode=True)[1]
    assert torch.equal(ref, alt)


# ---------------------------------------------------------------- reference vectors at the BASELINE shapes
FULL_CORR = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "fullsize_corr_*.npz")))
FULL_COORDS = {"grid": (0.0, 0.0), "half": (0.0, 0.5), "s3": (3.0, 0.37), "s20": (20.0, 0.0)}


@pytest.mark.parametrize("path", FULL_CORR, ids=[os.path.basename(p) for p in FULL_CORR])
def test_reference_vectors_at_baseline_shapes(path):
    """fullsize_corr_*.npz: the UNMODIFIED reference CorrBlock (CPU) at D=256 on 46x62 (configs 1/5) and 47x156
    (config 2) feature maps -- pyramid maps of 24 queries, lookups of 196 queries, four coordinate regimes.
    Inputs are regenerated from the seed.  Bars: volume <= 1e-3 relative (tensor-core operands) / 2e-6 (fp32 path),
    lookups <= 1e-5 of max|ref| on a pyramid that agrees with the reference's to fp32 rounding."""
    from weights import seeded_coords, seeded_fmaps

    m = ff()
    g = np.load(path)
    b, d, h, w, seed = [int(v) for v in g["shape"]]
    n = h * w
    f1, f2 = seeded_fmaps(seed, b, d, h, w)
    ql, q = torch.from_numpy(g["queries_levels"]).to(DEV), torch.from_numpy(g["queries"]).to(DEV)
    for prec, vtol, ltol in (("fp32", 2e-6, 1e-5), ("fp16", 1e-3, 1e-3), ("tf32", 1e-3, 1e-3), ("bf16x3", 3e-5, 3e-5)):
        blk = m.CorrBlock(t(f1), t(f2), num_levels=4, radius=4, precision=prec, sampler="cpu")
        for i in range(4):
            got = blk.corr_pyramid[i][ql, 0].cpu().numpy()
            assert got.shape == g[f"level{i}"].shape
            assert rel_fro(got, g[f"level{i}"]) <= vtol, (prec, i, rel_fro(got, g[f"level{i}"]))
        for k, (sigma, offset) in FULL_COORDS.items():
            c = seeded_coords(seed + 7, b, h, w, sigma, offset)
            out = blk(t(c)).view(b, 324, n).permute(0, 2, 1).reshape(b * n, 324)[q].cpu().numpy()
            err = np.abs(out - g[f"lookup_{k}"]).max() / float(g[f"lookup_{k}_absmax"])
            assert err <= ltol, (prec, k, err)
        del blk


E2E_FULL = os.path.join(os.path.dirname(__file__), "golden", "ffraft_e2e_full.npz")


@pytest.mark.parametrize("tag", ["c1", "c2", "c4"])
def test_e2e_epe_at_baseline_shapes(tag):
    """north_star: "end-point error after 12 refinement iterations within 0.01 px" -- against FF_RAFT_FUSION (the
    unmodified reference, CPU fp32) at 368x496 x12 (config 1), 376x1248 x12 (config 2) and 440x1024 x32 (config 4,
    evaluate.py:62/105 iterations), contractive flow head (0.2 px of motion per iteration, like a converged network).
    Host convolutions run in fp32 (TF32 off) so that the difference isolates the correlation path.
    The multi-pixel-motion case is test_lookups_along_the_reference_trajectory_with_large_motion: end to end it is
    chaotic -- the reference's own TF32 configuration drifts 9.5 px from its fp32 run there (oracle/make_golden.py)."""
    import sys

    sys.path.insert(0, os.path.dirname(__file__))
    from weights import DAMPED_GAIN, LIVELY_GAIN, fill_state_dict, synthetic_pair
    from focusflow_official_b200.host import FocusRAFT

    g = np.load(E2E_FULL)
    b, hh, ww, iters = [int(v) for v in g[f"{tag}_shape"]]
    gain = LIVELY_GAIN if str(g[f"{tag}_gain"]) == "lively" else DAMPED_GAIN
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    model = FocusRAFT()
    sd = model.state_dict()
    fill_state_dict(sd, seed=1234, flow_gain=gain)
    model.load_state_dict(sd)
    model = model.to(DEV).eval()
    im1, im2, m1, m2 = (x.to(DEV) for x in synthetic_pair(b, hh, ww, seed=4321))
    ref_lo = torch.from_numpy(g[f"{tag}_flow_lo"])
    ref_up = torch.from_numpy(g[f"{tag}_flow_up_s4"])
    drift_mean, drift_max = (float(v) for v in g[f"{tag}_ref_tf32_drift"])
    results = {}
    for prec in ("fp16", "fp32"):
        model.flow_net.corr_precision = prec
        with torch.no_grad():
            lo, up = model(im1, im2, m1, m2, raft_iters=iters, test_mode=True)
        epe = torch.linalg.norm(up.cpu()[:, :, ::4, ::4] - ref_up, dim=1)
        epe_lo = torch.linalg.norm(lo.cpu() - ref_lo, dim=1) * 8.0          # in full-resolution pixels
        results[prec] = (float(epe.mean()), float(epe.max()))
        print(f"e2e_full {tag} {prec}: EPE mean {float(epe.mean()):.2e} max {float(epe.max()):.2e} "
              f"(1/8-res flow x8: max {float(epe_lo.max()):.2e}); |flow| mean {float(ref_up.abs().mean()):.1f}; "
              f"reference TF32-vs-fp32 drift mean {drift_mean:.2e} max {drift_max:.2e}")
    # exact operands: the 0.01 px bar after 12 iterations (the north-star's statement); after 32 iterations even this path
    # -- which differs from the CPU reference only by cuDNN-vs-CPU convolution rounding -- sits at 1.2e-2 px max
    assert results["fp32"][1] <= (1e-2 if iters <= 12 else 2.5e-2), (tag, "fp32", results["fp32"])
    assert results["fp32"][0] <= 1e-3, (tag, "fp32", results["fp32"])
    # fp16 operands (the default; same 10-bit mantissa as the reference's TF32 GPU arithmetic): the 0.01 px bar at 12
    # iterations, and never further from the reference's fp32 run than 1.5x what the reference's OWN TF32 configuration
    # drifts from it (at 32 iterations that self-drift is 0.13 px: the refinement amplifies operand rounding)
    assert results["fp16"][1] <= max(1e-2, 1.5 * drift_max), (tag, "fp16", results["fp16"], drift_max)
    assert results["fp16"][0] <= max(2e-3, 1.5 * drift_mean), (tag, "fp16", results["fp16"], drift_mean)
    if iters <= 12:
        assert results["fp16"][1] <= 1e-2, (tag, "fp16", results["fp16"])


def test_lookups_along_the_reference_trajectory_with_large_motion():
    """the damped flow head keeps every lookup within ~1 px of the integer grid, so the end-to-end EPE says
    little about the correlation values.  Here the reference ran with the lively head (2.5 px of motion per iteration;
    |coords - grid| up to 15 cells at the end): its CorrBlock inputs `coords` at each of the 12 iterations and its
    lookup outputs for 64 queries are the fixture.  Our encoder (fp32 convolutions) produces the feature maps, our block
    is evaluated along the REFERENCE's trajectory: fp32 operands <= 2e-5 of max|ref| (the encoders differ by conv
    rounding), the tensor-core operand modes <= 1e-3 -- BASELINE's volume bar, since a lookup is a convex combination
    of volume entries -- for both output layouts."""
    import sys

    sys.path.insert(0, os.path.dirname(__file__))
    from weights import LIVELY_GAIN, fill_state_dict, synthetic_pair
    from focusflow_official_b200.host import FocusRAFT

    m = ff()
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ffraft_trajectory_c1_lively.npz"))
    b, hh, ww, iters = [int(v) for v in g["shape"]]
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    model = FocusRAFT()
    sd = model.state_dict()
    fill_state_dict(sd, seed=1234, flow_gain=LIVELY_GAIN)
    model.load_state_dict(sd)
    model = model.to(DEV).eval()
    im1, im2, m1, _ = (x.to(DEV) for x in synthetic_pair(b, hh, ww, seed=4321))
    from focusflow_official_b200.host.focusraft import init_point_mask

    mk1, mk2 = init_point_mask(m1, 3)
    sc = lambda x: 2 * (x / 255.0) - 1.0
    with torch.no_grad():
        fnet = model.flow_net.fnet
        fmap1, fmap2 = fnet(sc(im1), sc(mk1)).float(), fnet(sc(im2), sc(mk2)).float()
    q = torch.from_numpy(g["queries"]).to(DEV)
    n = (hh // 8) * (ww // 8)
    assert float(np.abs(g["coords"][-1] - g["coords"][0]).max()) > 10          # the trajectory really moves
    for prec, tol in (("fp32", 2e-5), ("fp16", 1e-3), ("tf32", 1e-3), ("bf16x3", 5e-5)):
        for cl in (False, True):
            blk = m.CorrBlock(fmap1, fmap2, precision=prec, sampler="cpu", channels_last=cl)
            worst = 0.0
            for it in range(iters):
                out = blk(t(g["coords"][it]))
                got = out.reshape(b, 324, n)[0].t()[q].cpu().numpy()
                worst = max(worst, float(np.abs(got - g["lookups"][it]).max() / g["absmax"][it]))
            print(f"trajectory {prec} channels_last={cl}: worst lookup error {worst:.2e} of max|ref|")
            assert worst <= tol, (prec, cl, worst)


# ---------------------------------------------------------------- opt-in half-precision pyramid storage
@pytest.mark.parametrize("shape,nl,radius", [((1, 256, 46, 62), 4, 4), ((2, 64, 17, 21), 4, 4), ((1, 64, 47, 156), 4, 4),
                                             ((2, 32, 24, 40), 3, 3), ((1, 40, 16, 24), 2, 2), ((2, 48, 33, 47), 4, 1)])
@pytest.mark.parametrize("precision", ["fp16", "tf32"])
def test_fp16_storage_build_and_lookup(shape, nl, radius, precision):
    """CorrBlock(storage="fp16") (ffcorr_build_tiled_f16 + ffcorr_lookup_tiled_f16):
      * every stored level == fp16(the fp32-stored level) exactly: same accumulators, same poolings, one rounding at the
        store (the fused fp32 build is the reference here, itself bit-identical to volume + avg_pool2d);
      * volume within 1e-3 of the exact fp64 contraction (BASELINE bar), pooled levels likewise;
      * lookups <= 1e-5 of max|ref| against the CPU oracle applied to the SAME fp16-rounded pyramid, every coordinate
        regime, both output layouts."""
    m = ff()
    b, d, h, w = shape
    rng = np.random.default_rng(77)
    f1 = (rng.standard_normal(shape) * 4.4).astype(np.float32)
    f2 = (rng.standard_normal(shape) * 4.4).astype(np.float32)
    blk32 = m.CorrBlock(t(f1), t(f2), num_levels=nl, radius=radius, precision=precision)
    blk16 = m.CorrBlock(t(f1), t(f2), num_levels=nl, radius=radius, precision=precision, storage="fp16", channels_last=True)
    assert blk16._levels[0].dtype == torch.float16 and blk16.storage == "fp16"
    assert sum(l.numel() * l.element_size() for l in blk16._levels) * 2 == sum(l.numel() * l.element_size() for l in blk32._levels)
    ref64 = co.volume_f64(f1, f2)
    pyr16 = []
    for i in range(nl):
        a32 = blk32.corr_pyramid[i]
        a16 = blk16.corr_pyramid[i]
        assert a16.dtype == torch.float32 and a16.shape == a32.shape
        assert torch.equal(a16, a32.half().float()), (i, float((a16 - a32.half().float()).abs().max()))
        pyr16.append(a16[:, 0].cpu().numpy())
    assert rel_fro(pyr16[0], ref64) <= 1e-3
    for name, c in _coords_cases(rng, b, h, w).items():
        ref = co.lookup(pyr16, c, radius)
        for cl in (True, False):
            blk16.channels_last = cl
            got = blk16(t(c))
            assert got.shape == ref.shape and got.dtype == torch.float32
            assert got.is_contiguous(memory_format=torch.channels_last) if cl else got.is_contiguous()
            assert max_rel(got.cpu().numpy(), ref) <= 1e-5, (name, cl, max_rel(got.cpu().numpy(), ref))


def test_fp16_storage_refuses_what_it_cannot_do():
    m = ff()
    f = torch.randn(1, 32, 16, 24, device=DEV)
    with pytest.raises(ValueError):
        m.CorrBlock(f, f, storage="fp16", precision="fp32")        # the exact path has no half-precision store
    with pytest.raises(ValueError):
        m.CorrBlock(f, f, storage="bf16")
    with pytest.raises(ValueError):
        m.CorrBlock(f, f, storage="fp16", num_levels=1)
    # tile_levels(storage="fp16") + lookup_tiled reproduce the oracle on the rounded levels too
    lv = [torch.randn(16 * 24, 1, 16 >> i, 24 >> i, device=DEV) for i in range(3)]
    c = m.coords_grid(1, 16, 24, DEV) + 0.3
    got = m.lookup_tiled(m.tile_levels(lv, storage="fp16"), c, 3, channels_last=True)
    ref = m.lookup([x.half().float() for x in lv], c, 3)
    assert float((got - ref).abs().max()) <= 1e-5 * float(ref.abs().max())


def test_fp16_storage_end_to_end_and_full_size():
    """EPE of the host model with the half-precision pyramid against the reference goldens (config 1 and config 2
    shapes, 12 iterations), and the config-2 batch-8 build + lookup at full size against the fp32-stored block."""
    import sys

    sys.path.insert(0, os.path.dirname(__file__))
    from weights import fill_state_dict, synthetic_pair
    from focusflow_official_b200.host import FocusRAFT

    g = np.load(E2E_FULL)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    model = FocusRAFT()
    sd = model.state_dict()
    fill_state_dict(sd, seed=1234)
    model.load_state_dict(sd)
    model = model.to(DEV).eval()
    model.flow_net.corr_storage = "fp16"
    for tag in ("c1", "c2"):
        b, hh, ww, iters = [int(v) for v in g[f"{tag}_shape"]]
        im1, im2, m1, m2 = (x.to(DEV) for x in synthetic_pair(b, hh, ww, seed=4321))
        with torch.no_grad():
            lo, up = model(im1, im2, m1, m2, raft_iters=iters, test_mode=True)
        epe = torch.linalg.norm(up.cpu()[:, :, ::4, ::4] - torch.from_numpy(g[f"{tag}_flow_up_s4"]), dim=1)
        print(f"e2e_full {tag} fp16 operands + fp16 storage: EPE mean {float(epe.mean()):.2e} max {float(epe.max()):.2e}")
        # MEASURED: 1.5e-2 px max / 2.6e-3 mean at config 1, 4.2e-2 / 2.7e-3 at config 2 -- one fp16 rounding of every stored correlation value (up to
        # 0.03 absolute at |corr| ~ 100) costs more accuracy than rounding the GEMM operands, whose errors average out
        # over the 256-term dot product.  Half-precision storage therefore does NOT meet the north-star's 0.01 px bar
        # and stays opt-in; this assertion only bounds the damage.
        assert float(epe.max()) <= 8e-2 and float(epe.mean()) <= 5e-3, (tag, float(epe.max()), float(epe.mean()))
    m = ff()
    torch.manual_seed(5)
    f1 = torch.randn(8, 256, 47, 156, device=DEV) * 4.4
    f2 = torch.randn(8, 256, 47, 156, device=DEV) * 4.4
    coords = m.coords_grid(8, 47, 156, DEV) + torch.randn(8, 2, 47, 156, device=DEV) * 3
    a = m.CorrBlock(f1, f2, channels_last=True)(coords)
    bq = m.CorrBlock(f1, f2, channels_last=True, storage="fp16")(coords)
    assert float((a - bq).abs().max()) <= 1e-3 * float(a.abs().max())      # one fp16 rounding of every stored value
    assert float((a - bq).norm() / a.norm()) <= 3e-4


@pytest.mark.parametrize("scale1,scale2", [(1e5, 1e5), (1e-6, 1e-6), (3e7, 2e-7), (1e15, 1e-3), (4.4, 4.4)])
@pytest.mark.parametrize("layout", ["tiled", "rowmajor"])
def test_fp16_operands_are_block_scaled_into_range(scale1, scale2, layout):
    """fp16 operands used to need |fmap| < 65504 (1e5 became inf, 1e-6 flushed to zero).  The
    pre-pass now scales every feature map by an exact power of two into fp16's range (one exponent per batch item and
    operand, from max|fmap|) and the GEMM epilogue undoes it: the volume keeps the 1e-3 bar for activations of any
    magnitude, also when the two maps -- or the items of a batch -- differ by 20 orders of magnitude."""
    m = ff()
    rng = np.random.default_rng(3)
    b, d, h, w = 2, 64, 16, 24
    f1 = (rng.standard_normal((b, d, h, w)) * scale1).astype(np.float32)
    f2 = (rng.standard_normal((b, d, h, w)) * scale2).astype(np.float32)
    f1[1] *= np.float32(1e-3)                      # per-item exponents: the second pair lives 3 decades lower
    ref = co.volume_f64(f1, f2)
    blk = m.CorrBlock(t(f1), t(f2), precision="fp16", layout=layout)
    got = blk.corr_pyramid[0][:, 0].cpu().numpy().astype(np.float64)
    assert np.isfinite(got).all()
    n = h * w
    for item in range(b):
        sl = slice(item * n, (item + 1) * n)
        assert rel_fro(got[sl], ref[sl]) <= 1e-3, (item, rel_fro(got[sl], ref[sl]))
    # and the pooled levels / a lookup stay finite and consistent with the volume
    out = blk(m.coords_grid(b, h, w, DEV) + 0.4)
    assert torch.isfinite(out).all()


# ---------------------------------------------------------------- FlowFormer operations against the reference's own functions
@pytest.mark.parametrize("tag", ["a", "b"])
def test_flowformer_ops_against_reference_golden(tag):
    """tests/golden/flowformer_ops.npz = outputs of the UNMODIFIED MemoryEncoder.corr (encoder.py:337-348),
    MemoryDecoder.encode_flow_token (decoder.py:185-203) and ReverseCostExtractor.forward (decoder.py:119-149), run on CPU
    by oracle/make_golden_flowformer.py.  cost volume <= 1e-3 (tensor-core operands) / 2e-6 (fp32); the two lookups
    <= 1e-5 of max|ref| on the reference's own cost maps."""
    from focusflow_official_b200 import flowformer as fl

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "flowformer_ops.npz"))
    b, dim, h, w, heads = [int(v) for v in g[f"{tag}_shape"]]
    f1, f2 = t(g[f"{tag}_fmap1"]), t(g[f"{tag}_fmap2"])
    ref_cost = g[f"{tag}_cost"]
    for prec, tol in (("fp16", 1e-3), ("tf32", 1e-3), ("fp32", 2e-6)):
        got = fl.cost_volume(f1, f2, heads=heads, precision=prec).cpu().numpy()
        assert got.shape == ref_cost.shape and rel_fro(got, ref_cost) <= tol, (prec, rel_fro(got, ref_cost))
    cost_maps = t(ref_cost).permute(0, 2, 3, 1, 4, 5).reshape(b * h * w, heads, h, w).contiguous()
    c0, c1 = t(g[f"{tag}_coords0"]), t(g[f"{tag}_coords1"])
    tok = fl.encode_flow_token(cost_maps, c1).cpu().numpy()
    assert tok.shape == g[f"{tag}_flow_token"].shape and max_rel(tok, g[f"{tag}_flow_token"]) <= 1e-5
    rev = fl.reverse_cost_tokens(cost_maps, c0, c1).cpu().numpy()
    assert rev.shape == g[f"{tag}_reverse"].shape
    assert max_rel(rev, g[f"{tag}_reverse"]) <= 1e-5, max_rel(rev, g[f"{tag}_reverse"])


# ---------------------------------------------------------------- channels-last (NHWC) lookups: SURVEY 8b / 8f N3
@pytest.mark.parametrize("shape,radius,nl", [((1, 46, 62), 4, 4), ((2, 17, 21), 4, 4), ((1, 24, 40), 3, 4), ((2, 12, 9), 2, 3),
                                             ((1, 9, 33), 1, 2), ((3, 8, 8), 4, 1), ((1, 47, 156), 4, 4), ((2, 13, 11), 1, 1)])
def test_channels_last_lookup_is_the_same_values_in_nhwc_memory(shape, radius, nl):
    """out_channels_last = 1 (include/ffcorr.h): [B, h, w, L*K*K] storage, the same values, for the row-major
    kernel, the tiled kernel (8 queries x all levels per warp, TMA bulk store) and the chunked (AlternateCorrBlock)
    entry point; query counts that are not multiples of 8 / 32, 1-4 levels, radii 1-4, every coordinate regime."""
    m = ff()
    b, h, w = shape
    rng = np.random.default_rng(5)
    q = b * h * w
    k2 = (2 * radius + 1) ** 2
    pyr = [rng.standard_normal((q, h >> i, w >> i)).astype(np.float32) for i in range(nl)]
    levels = [t(p[:, None]) for p in pyr]
    tl = m.tile_levels(levels)
    for name, c in _coords_cases(rng, b, h, w).items():
        cd = t(c)
        ref = co.lookup(pyr, c, radius)
        for fn, lv in ((m.lookup, levels), (m.lookup_tiled, tl)):
            nchw = fn(lv, cd, radius)
            nhwc = fn(lv, cd, radius, channels_last=True)
            assert nhwc.shape == (b, nl * k2, h, w) and nhwc.is_contiguous(memory_format=torch.channels_last) or nl * k2 == 1
            assert nhwc.permute(0, 2, 3, 1).is_contiguous()
            # same values: bit-identical per evaluation path; a warp takes the exact per-tap path when ANY of its 32
            # windows sits on an integer coordinate, and the two kernels group windows differently (32 queries of one
            # level vs 8 queries x all levels), so single windows may differ by the ~1e-7 between the two paths
            scale = max(float(nchw.abs().max()), 1e-20)
            assert float((nhwc - nchw).abs().max()) <= 2e-6 * scale, (fn.__name__, name)
            if nl == 1:
                assert torch.equal(nhwc.contiguous(), nchw), (fn.__name__, name)
            assert max_rel(nhwc.cpu().numpy(), ref) <= 1e-5, (fn.__name__, name)


def test_channels_last_blocks_and_the_host_model_use_no_layout_copy():
    m = ff()
    torch.manual_seed(3)
    f1, f2 = torch.randn(2, 64, 24, 40, device=DEV), torch.randn(2, 64, 24, 40, device=DEV)
    coords = m.coords_grid(2, 24, 40, DEV) + torch.randn(2, 2, 24, 40, device=DEV) * 2
    a = m.CorrBlock(f1, f2)(coords)
    for blk in (m.CorrBlock(f1, f2, channels_last=True), m.AlternateCorrBlock(f1, f2, channels_last=True, chunk=200),
                m.CorrBlock(f1, f2, channels_last=True, precision="fp32")):
        out = blk(coords)
        assert out.is_contiguous(memory_format=torch.channels_last)
        assert out.contiguous(memory_format=torch.channels_last).data_ptr() == out.data_ptr()     # no copy for convc1
        if blk.__class__.__name__ != "CorrBlock" or blk._tiled:
            assert float((out - a).abs().max()) <= 2e-6 * float(a.abs().max())
    # autograd through a channels-last block == through the NCHW one
    g = torch.randn_like(a)
    grads = []
    for cl in (False, True):
        x1, x2 = f1.clone().requires_grad_(True), f2.clone().requires_grad_(True)
        out = m.CorrBlock(x1, x2, channels_last=cl)(coords)
        (out * g).sum().backward()
        grads.append((x1.grad, x2.grad))
    for ga, gb in zip(grads[0], grads[1]):
        assert float((ga - gb).abs().max()) <= 1e-5 * float(ga.abs().max())


def test_context_encoder_in_channels_last_gives_the_same_flow():
    """bench.py runs the context encoder (eval-mode BatchNorm) in channels_last: the same convolutions without layout
    transposes around them.  Flow must agree with the NCHW run to well inside what TF32 host convolutions move it."""
    from test_host_model import make_model
    from weights import synthetic_pair

    model = make_model().to(DEV).eval()
    model.flow_net.update_block.to(memory_format=torch.channels_last)
    model.flow_net.update_channels_last = True
    im1, im2, m1, _ = (x.to(DEV) for x in synthetic_pair(2, 128, 192, seed=5))
    before = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        with torch.no_grad():
            torch.backends.cudnn.allow_tf32 = False
            exact = model(im1, im2, m1, None, raft_iters=12, test_mode=True)[1]
            torch.backends.cudnn.allow_tf32 = True                         # the reference's run setting (common.py:25-27)
            ref = model(im1, im2, m1, None, raft_iters=12, test_mode=True)[1]
            model.flow_net.cnet.to(memory_format=torch.channels_last)
            got = model(im1, im2, m1, None, raft_iters=12, test_mode=True)[1]
            torch.backends.cudnn.allow_tf32 = False
            got_exact = model(im1, im2, m1, None, raft_iters=12, test_mode=True)[1]
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = before
    drift = (ref - exact).norm(dim=1).max().item()                         # what TF32 alone does to the flow
    epe = (got - ref).norm(dim=1).max().item()
    epe_exact = (got_exact - exact).norm(dim=1).max().item()               # fp32 convolutions: only the summation order differs
    print(f"context encoder NHWC vs NCHW: EPE max {epe:.2e} (TF32), {epe_exact:.2e} (fp32); TF32-vs-fp32 drift {drift:.2e}")
    assert torch.isfinite(got).all()
    assert epe <= max(2.0 * drift, 2e-3)
    assert epe_exact <= 1e-3


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_launches_follow_the_tensors_device_not_the_current_one():
    """with cuda:0 current and the model on cuda:1, kernels, TMA descriptors and the stream must be those of
    cuda:1 (ATen guards the device the same way); mixing devices in one call is a ValueError."""
    m = ff()
    assert torch.cuda.current_device() == 0
    dev1 = torch.device("cuda", 1)
    torch.manual_seed(9)
    f1, f2 = torch.randn(1, 64, 24, 32), torch.randn(1, 64, 24, 32)
    coords = m.coords_grid(1, 24, 32, "cpu") + torch.randn(1, 2, 24, 32)
    ref = m.CorrBlock(f1.to(DEV), f2.to(DEV))(coords.to(DEV))
    side = torch.cuda.Stream(device=dev1)
    with torch.cuda.stream(side):                       # current stream of cuda:1 only; current device stays 0
        pass
    blk = m.CorrBlock(f1.to(dev1), f2.to(dev1))
    out = blk(coords.to(dev1))
    assert out.device == dev1 and torch.cuda.current_device() == 0
    assert torch.equal(out.cpu(), ref.cpu())
    cv = m.FunctionCorrelation(f1.to(dev1), f2.to(dev1))
    assert torch.equal(cv.cpu(), m.FunctionCorrelation(f1.to(DEV), f2.to(DEV)).cpu())
    x1 = f1.to(dev1).requires_grad_(True)
    m.CorrBlock(x1, f2.to(dev1))(coords.to(dev1)).sum().backward()
    assert x1.grad is not None and x1.grad.device == dev1 and torch.isfinite(x1.grad).all()
    with pytest.raises(ValueError):
        m.CorrBlock(f1.to(DEV), f2.to(dev1))
    with pytest.raises(ValueError):
        blk(coords.to(DEV))


# ---------------------------------------------------------------- grouped ("G32") tile order
@pytest.mark.parametrize("shape,nl,radius", [((1, 256, 46, 62), 4, 4), ((2, 64, 17, 21), 4, 4), ((1, 64, 47, 156), 4, 4),
                                             ((2, 32, 24, 40), 3, 3), ((1, 40, 16, 24), 2, 2), ((3, 48, 33, 47), 4, 1)])
@pytest.mark.parametrize("precision", ["fp16", "tf32"])
def test_grouped_layout_is_the_tiled_layout_reordered(shape, nl, radius, precision):
    """CorrBlock(layout="grouped"): [group of 32 queries][tile][query][4][4] storage written by the fused build's epilogue
    in contiguous 8 KB / 4 KB pieces.  Same GEMM, same pooling, same lookup code -- only addresses differ -- so every level
    and every lookup (both output layouts, every coordinate regime, query counts that are not multiples of 32) must be
    BIT-identical to the tiled block's."""
    m = ff()
    b, d, h, w = shape
    rng = np.random.default_rng(123)
    f1 = t((rng.standard_normal(shape) * 4.4).astype(np.float32))
    f2 = t((rng.standard_normal(shape) * 4.4).astype(np.float32))
    ref = m.CorrBlock(f1, f2, num_levels=nl, radius=radius, precision=precision, layout="tiled")
    for cl in (False, True):
        blk = m.CorrBlock(f1, f2, num_levels=nl, radius=radius, precision=precision, layout="grouped", channels_last=cl)
        assert blk._levels[0].dim() == 3 and blk._levels[0].shape[0] == b * ((h * w + 31) // 32)
        ref.channels_last = cl
        for i in range(nl):
            assert torch.equal(blk.corr_pyramid[i], ref.corr_pyramid[i]), (cl, i)
        for name, c in _coords_cases(rng, b, h, w).items():
            got, want = blk(t(c)), ref(t(c))
            assert got.shape == want.shape and torch.equal(got, want), (cl, name, int((got != want).sum()))
