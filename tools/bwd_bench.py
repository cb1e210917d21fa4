"""Training-path timings: forward + backward of CorrBlock at config 5 (B=8, 368x496), this repo's
kernels vs stock PyTorch autograd of the reference's op sequence on the same GPU.  Measurement only."""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import focusflow_official_b200 as ff  # noqa: E402

dev = "cuda:0"
b, d, h, w = 8, 256, 46, 62
iters = 12
n = h * w
torch.manual_seed(0)
f1 = (torch.randn(b, d, h, w, device=dev) * 4.4).requires_grad_(True)
f2 = (torch.randn(b, d, h, w, device=dev) * 4.4).requires_grad_(True)
coords = [ff.coords_grid(b, h, w, dev) + torch.randn(b, 2, h, w, device=dev) * 3 for _ in range(iters)]
gout = torch.randn(b, 324, h, w, device=dev)


def ours(precision):
    blk = ff.CorrBlock(f1, f2, precision=precision)
    loss = sum((blk(c) * gout).sum() for c in coords)
    loss.backward()


def stock():
    corr = torch.matmul(f1.view(b, d, n).transpose(1, 2), f2.view(b, d, n)).view(b, h, w, 1, h, w)
    corr = (corr / torch.sqrt(torch.tensor(d).float())).reshape(b * n, 1, h, w)
    pyr = [corr]
    for _ in range(3):
        corr = F.avg_pool2d(corr, 2, stride=2)
        pyr.append(corr)
    loss = 0
    r = 4
    for c0 in coords:
        c = c0.permute(0, 2, 3, 1)
        out = []
        for i, cr in enumerate(pyr):
            dx = torch.linspace(-r, r, 2 * r + 1, device=dev)
            delta = torch.stack(torch.meshgrid(dx, dx, indexing="ij"), axis=-1)
            cl = c.reshape(b * n, 1, 1, 2) / 2 ** i + delta.view(1, 9, 9, 2)
            hh, ww = cr.shape[-2:]
            xg, yg = cl.split([1, 1], dim=-1)
            grid = torch.cat([2 * xg / (ww - 1) - 1, 2 * yg / (hh - 1) - 1], dim=-1)
            out.append(F.grid_sample(cr, grid, align_corners=True).view(b, h, w, -1))
        o = torch.cat(out, dim=-1).permute(0, 3, 1, 2).contiguous().float()
        loss = loss + (o * gout).sum()
    loss.backward()


def timeit(fn, reps=5):
    for _ in range(2):
        f1.grad = f2.grad = None
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f1.grad = f2.grad = None
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


torch.backends.cuda.matmul.allow_tf32 = True
res = {"config": 5, "iters": iters, "stock_aten_tf32_ms": round(timeit(stock), 3)}
for p in ("fp16", "fp32"):
    res[f"ffcorr_{p}_ms"] = round(timeit(lambda: ours(p)), 3)
print(json.dumps(res))
if "--pwc" in sys.argv:
    # PWC cost volume forward + backward at the five config-3 level shapes
    for c, hh, ww in [(32, 112, 256), (64, 56, 128), (96, 28, 64), (128, 14, 32), (196, 7, 16)]:
        one = torch.randn(16, c, hh, ww, device=dev, requires_grad=True)
        two = torch.randn(16, c, hh, ww, device=dev, requires_grad=True)
        go = torch.randn(16, 81, hh, ww, device=dev)

        def fb():
            one.grad = two.grad = None
            ff.FunctionCorrelation(one, two).backward(go)

        def fwd():
            with torch.no_grad():
                ff.FunctionCorrelation(one, two)

        for _ in range(2):
            fb()
        torch.cuda.synchronize()
        ts = []
        for fn in (fwd, fb):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / 5)
        print(json.dumps({"kernel": "pwc81 fwd / fwd+bwd", "C": c, "H": hh, "W": ww, "fwd_ms": round(ts[0], 4),
                          "fwd_bwd_ms": round(ts[1], 4)}))
if "--profile" in sys.argv:
    from torch.profiler import profile, ProfilerActivity
    f1.grad = f2.grad = None
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        ours("fp16")
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
