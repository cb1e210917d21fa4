"""Per-kernel timing of the correlation path (CUDA events, L2 flushed between launches).

Prints one JSON line per kernel with the algorithmic bytes / flops of SURVEY.md 8d and the
achieved fraction of the measured peaks in MEASURED_PEAKS.json.  Development tool; bench.py
reuses `time_kernels()` for its roofline block.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        d["source"] = "measured"
        return d
    d = dict(FALLBACK_PEAKS)
    d["source"] = "fallback"
    return d


class L2Flusher:
    def __init__(self, dev, mb=512):
        self.buf = torch.empty(mb * 1024 * 1024, dtype=torch.uint8, device=dev)

    def __call__(self):
        self.buf.zero_()


def time_cuda(fn, iters=10, warmup=3, flush=None):
    """Median / min duration in ms of fn() measured with CUDA events on the current stream."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def raft_shapes(config):
    return {1: (1, 256, 46, 62), 2: (8, 256, 47, 156), 4: (8, 256, 55, 128), 5: (8, 256, 46, 62),
            6: (8, 256, 56, 128), 7: (8, 256, 48, 160)}[config]  # 6/7: tile-aligned shapes (diagnosis)


def time_kernels(config=2, iters=10, precision="fp16", dev="cuda:0", sigma=3.0, verbose=False, warmup=3, smooth=False,
                 only=None):
    import focusflow_official_b200 as ff
    from focusflow_official_b200 import _lib

    peaks = load_peaks()
    L = _lib.lib()
    if os.environ.get("FFCORR_L2_FETCH"):
        _lib.check(L.ffcorr_set_l2_fetch_granularity(int(os.environ["FFCORR_L2_FETCH"])), "l2 fetch")
    b, d, h, w = raft_shapes(config)
    n = h * w
    torch.manual_seed(1234)
    f1 = torch.randn(b, d, h, w, device=dev) * 4.4
    f2 = torch.randn(b, d, h, w, device=dev) * 4.4
    nl, r = 4, 4
    levels = [torch.empty(b * n, 1, h >> i, w >> i, device=dev) for i in range(nl)]
    code = _lib.PRECISIONS[precision]
    ws_bytes = L.ffcorr_volume_workspace_bytes(b, d, h, w, code)
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
    ptrs = _lib.ptr_array(levels)
    stream = _lib.current_stream()
    if smooth:
        # a flow field like the model produces: a low-frequency displacement plus sub-pixel jitter
        ys = torch.linspace(0, 3.14159, h, device=dev).view(1, 1, h, 1)
        xs = torch.linspace(0, 6.28318, w, device=dev).view(1, 1, 1, w)
        flow = torch.cat([5.3 * torch.sin(xs + ys), 2.7 * torch.cos(xs - ys)], 1).expand(b, 2, h, w)
        coords = ff.coords_grid(b, h, w, dev) + flow + torch.randn(b, 2, h, w, device=dev) * 0.25
    else:
        coords = ff.coords_grid(b, h, w, dev) + torch.randn(b, 2, h, w, device=dev) * sigma
    out = torch.empty(b, nl * 81, h, w, device=dev)
    flush = L2Flusher(dev)

    tl = [torch.empty(b * n, int(L.ffcorr_tiled_map_elems(h, w, i)), device=dev) for i in range(nl)]
    tptrs = _lib.ptr_array(tl)
    tiled_ok = bool(L.ffcorr_tiled_supported(nl, h, w)) and code != _lib.PREC_FP32

    def k_volume():
        _lib.check(L.ffcorr_volume_f32(f1.data_ptr(), f2.data_ptr(), levels[0].data_ptr(), b, d, h, w, code,
                                       ws.data_ptr(), ws_bytes, stream), "volume")

    def k_pyramid():
        _lib.check(L.ffcorr_pyramid_f32(ptrs, nl, b * n, h, w, stream), "pyramid")

    def k_lookup():
        _lib.check(L.ffcorr_lookup_f32(ptrs, nl, coords.data_ptr(), out.data_ptr(), b, h, w, r, 1, 0, stream), "lookup")

    def k_volume_t():
        _lib.check(L.ffcorr_volume_tiled_f32(f1.data_ptr(), f2.data_ptr(), tl[0].data_ptr(), b, d, h, w, code,
                                             ws.data_ptr(), ws_bytes, stream), "volume_tiled")

    def k_pyramid_t():
        _lib.check(L.ffcorr_pyramid_tiled_f32(tptrs, nl, b * n, h, w, stream), "pyramid_tiled")

    def k_build_t():
        _lib.check(L.ffcorr_build_tiled_f32(f1.data_ptr(), f2.data_ptr(), tptrs, nl, b, d, h, w, code, ws.data_ptr(), ws_bytes,
                                            stream), "build_tiled")

    def k_lookup_t():
        _lib.check(L.ffcorr_lookup_tiled_f32(tptrs, nl, coords.data_ptr(), out.data_ptr(), b, h, w, r, 1, 0, stream), "lookup_tiled")

    out_nhwc = torch.empty(b, h, w, nl * 81, device=dev)

    def k_lookup_t_nhwc():
        _lib.check(L.ffcorr_lookup_tiled_f32(tptrs, nl, coords.data_ptr(), out_nhwc.data_ptr(), b, h, w, r, 1, 1, stream), "lookup_tiled_nhwc")

    th = [torch.empty(b * n, int(L.ffcorr_tiled_map_elems(h, w, i)), device=dev, dtype=torch.float16) for i in range(nl)]
    hptrs = _lib.ptr_array(th)

    def k_build_h():
        _lib.check(L.ffcorr_build_tiled_f16(f1.data_ptr(), f2.data_ptr(), hptrs, nl, b, d, h, w, code, ws.data_ptr(), ws_bytes,
                                            stream), "build_tiled_f16")

    def k_lookup_h():
        _lib.check(L.ffcorr_lookup_tiled_f16(hptrs, nl, coords.data_ptr(), out_nhwc.data_ptr(), b, h, w, r, 1, 1, stream), "lookup_tiled_f16")

    ng = (n + 31) // 32
    tg = [torch.empty(int(L.ffcorr_grouped_level_elems(h, w, i, b, n)), device=dev) for i in range(nl)]
    gptrs = _lib.ptr_array(tg)

    def k_build_g():
        _lib.check(L.ffcorr_build_grouped_f32(f1.data_ptr(), f2.data_ptr(), gptrs, nl, b, d, h, w, code, ws.data_ptr(), ws_bytes,
                                              stream), "build_grouped")

    def k_lookup_g():
        _lib.check(L.ffcorr_lookup_grouped_f32(gptrs, nl, coords.data_ptr(), out.data_ptr(), b, h, w, r, 1, 0, stream), "lookup_grouped")

    def k_lookup_g_nhwc():
        _lib.check(L.ffcorr_lookup_grouped_f32(gptrs, nl, coords.data_ptr(), out_nhwc.data_ptr(), b, h, w, r, 1, 1, stream), "lookup_grouped_nhwc")

    res = []
    lv_elems = [(h >> i) * (w >> i) for i in range(nl)]
    vol_flops = 2.0 * b * n * n * d
    vol_bytes = b * (2 * n * d * 4 + n * n * 4)
    pyr_bytes = 4.0 * b * n * sum(lv_elems)
    look_bytes = b * n * (nl * (2 * r + 2) ** 2 * 4 + nl * (2 * r + 1) ** 2 * 4 + 8)
    todo = [("volume", k_volume, vol_bytes, vol_flops), ("pyramid", k_pyramid, pyr_bytes, 0.0), ("lookup", k_lookup, look_bytes, 0.0)]
    if tiled_ok:  # same algorithmic bytes: the padding of the tiled layout is overhead, not work
        todo += [("volume_tiled", k_volume_t, vol_bytes, vol_flops), ("pyramid_tiled", k_pyramid_t, pyr_bytes, 0.0),
                 ("build_fused", k_build_t, vol_bytes + pyr_bytes - 4.0 * b * n * lv_elems[0], vol_flops),
                 ("lookup_tiled", k_lookup_t, look_bytes, 0.0), ("lookup_tiled_nhwc", k_lookup_t_nhwc, look_bytes, 0.0)]
        # half-precision storage: algorithmic bytes with 2-byte pyramid elements
        vol_h = b * (2 * n * d * 4 + n * n * 2)
        pyr_h = 2.0 * b * n * sum(lv_elems[1:])
        look_h = b * n * (nl * (2 * r + 2) ** 2 * 2 + nl * (2 * r + 1) ** 2 * 4 + 8)
        todo += [("build_fused_f16", k_build_h, vol_h + pyr_h, vol_flops), ("lookup_tiled_f16", k_lookup_h, look_h, 0.0)]
        todo += [("build_grouped", k_build_g, vol_bytes + pyr_bytes - 4.0 * b * n * lv_elems[0], vol_flops),
                 ("lookup_grouped", k_lookup_g, look_bytes, 0.0), ("lookup_grouped_nhwc", k_lookup_g_nhwc, look_bytes, 0.0)]
    if tiled_ok and (not only or "alt_lookup" in only):
        # memory-bounded AlternateCorrBlock: the pyramid of a 512 MiB query chunk is rebuilt for every lookup
        alt = ff.AlternateCorrBlock(f1, f2, num_levels=nl, radius=r, precision=precision)
        med, best = time_cuda(lambda: alt(coords), iters=iters, warmup=warmup, flush=flush)
        rec = {"kernel": "alt_lookup", "config": config, "ms": round(med, 4), "ms_min": round(best, 4), "chunk_queries": alt.chunk,
               "pyramid_buffer_MiB": round(sum(l.numel() for l in alt._levels) * 4 / 2 ** 20, 1), "precision": precision}
        res.append(rec)
        if verbose:
            print(json.dumps(rec), flush=True)
        del alt
    for name, fn, byts, flops in todo:
        if only and name not in only:
            # run (untimed) only what a selected kernel reads
            needs = {"pyramid": ["volume"], "lookup": ["volume", "pyramid"], "pyramid_tiled": ["volume_tiled"],
                     "lookup_tiled": ["build_fused"], "lookup_tiled_nhwc": ["build_fused"], "lookup_tiled_f16": ["build_fused_f16"],
                     "lookup_grouped": ["build_grouped"], "lookup_grouped_nhwc": ["build_grouped"]}
            if any(name in needs.get(o, []) for o in only):
                fn()
            continue
        med, best = time_cuda(fn, iters=iters, warmup=warmup, flush=flush)
        gbs = byts / (med * 1e-3) / 1e9
        rec = {"kernel": name, "config": config, "ms": round(med, 4), "ms_min": round(best, 4),
               "algorithmic_bytes": int(byts), "GBps": round(gbs, 1), "frac_hbm": round(gbs / peaks["hbm_gbs"], 4),
               "peaks": peaks["source"]}
        if flops:
            tf = flops / (med * 1e-3) / 1e12
            rec.update({"TFLOPs": round(tf, 1), "frac_bf16_burst": round(tf / peaks["bf16_tflops"], 4), "precision": precision})
        res.append(rec)
        if verbose:
            print(json.dumps(rec), flush=True)
    return res


def time_stock_aten(config=2, iters=5, dev="cuda:0", sigma=3.0, verbose=False, tf32=True):
    """The same operations through stock PyTorch CUDA ops (what the reference's corr.py / utils.py run on a GPU):
    torch.matmul + divide, 3x avg_pool2d, and per level meshgrid + grid_sample.  Measurement only."""
    import torch.nn.functional as F

    b, d, h, w = raft_shapes(config)
    n = h * w
    torch.manual_seed(1234)
    torch.backends.cuda.matmul.allow_tf32 = tf32
    f1 = torch.randn(b, d, h, w, device=dev) * 4.4
    f2 = torch.randn(b, d, h, w, device=dev) * 4.4
    coords = torch.stack(torch.meshgrid(torch.arange(h, device=dev), torch.arange(w, device=dev), indexing="ij")[::-1], 0).float()
    coords = coords[None].repeat(b, 1, 1, 1) + torch.randn(b, 2, h, w, device=dev) * sigma
    flush = L2Flusher(dev)
    state = {}

    def build():
        corr = torch.matmul(f1.view(b, d, n).transpose(1, 2), f2.view(b, d, n)).view(b, h, w, 1, h, w)
        corr = (corr / torch.sqrt(torch.tensor(d).float())).reshape(b * n, 1, h, w)
        pyr = [corr]
        for _ in range(3):
            corr = F.avg_pool2d(corr, 2, stride=2)
            pyr.append(corr)
        state["pyr"] = pyr

    def lookup():
        r = 4
        c = coords.permute(0, 2, 3, 1)
        out = []
        for i, corr in enumerate(state["pyr"]):
            dx = torch.linspace(-r, r, 2 * r + 1, device=dev)
            dy = torch.linspace(-r, r, 2 * r + 1, device=dev)
            delta = torch.stack(torch.meshgrid(dy, dx, indexing="ij"), axis=-1)
            cl = c.reshape(b * n, 1, 1, 2) / 2 ** i + delta.view(1, 2 * r + 1, 2 * r + 1, 2)
            hh, ww = corr.shape[-2:]
            xg, yg = cl.split([1, 1], dim=-1)
            grid = torch.cat([2 * xg / (ww - 1) - 1, 2 * yg / (hh - 1) - 1], dim=-1)
            out.append(F.grid_sample(corr, grid, align_corners=True).view(b, h, w, -1))
        return torch.cat(out, dim=-1).permute(0, 3, 1, 2).contiguous().float()

    res = []
    for name, fn in (("stock_aten_build", build), ("stock_aten_lookup", lookup)):
        med, best = time_cuda(fn, iters=iters, warmup=2, flush=flush)
        rec = {"kernel": name, "config": config, "ms": round(med, 4), "ms_min": round(best, 4), "tf32_matmul": tf32}
        res.append(rec)
        if verbose:
            print(json.dumps(rec), flush=True)
    return res


def time_pwc(iters=10, dev="cuda:0", batch=16, verbose=False, warmup=3):
    import focusflow_official_b200 as ff

    peaks = load_peaks()
    flush = L2Flusher(dev)
    shapes = [(32, 112, 256), (64, 56, 128), (96, 28, 64), (128, 14, 32), (196, 7, 16)]
    res = []
    tot_ms = tot_b = 0.0
    for c, hh, ww in shapes:
        one = torch.randn(batch, c, hh, ww, device=dev)
        two = torch.randn(batch, c, hh, ww, device=dev)
        med, best = time_cuda(lambda: ff.FunctionCorrelation(one, two), iters=iters, warmup=warmup, flush=flush)
        # back-to-back burst (no flush): what a level costs inside the decoder, where its inputs were just
        # produced; for the small levels the flushed single-launch number is dominated by cold misses
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        flush()
        e0.record()
        for _ in range(20):
            ff.FunctionCorrelation(one, two)
        e1.record()
        e1.synchronize()
        burst = e0.elapsed_time(e1) / 20
        byts = 4.0 * batch * hh * ww * (2 * c + 81)
        flops = 162.0 * c * batch * hh * ww
        tot_ms += med
        tot_b += byts
        rec = {"kernel": "pwc81", "C": c, "H": hh, "W": ww, "B": batch, "ms": round(med, 4), "ms_min": round(best, 4), "ms_burst": round(burst, 4),
               "GBps": round(byts / med / 1e6, 1), "frac_hbm": round(byts / med / 1e6 / peaks["hbm_gbs"], 4),
               "TFLOPs_fp32": round(flops / med / 1e9, 2)}
        res.append(rec)
        if verbose:
            print(json.dumps(rec), flush=True)
    rec = {"kernel": "pwc81_total", "ms": round(tot_ms, 4), "GBps": round(tot_b / tot_ms / 1e6, 1),
           "frac_hbm": round(tot_b / tot_ms / 1e6 / peaks["hbm_gbs"], 4)}
    res.append(rec)
    if verbose:
        print(json.dumps(rec), flush=True)
    return res


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--precision", default="fp16")
    ap.add_argument("--sigma", type=float, default=3.0)
    ap.add_argument("--smooth", action="store_true", help="smooth synthetic flow instead of iid noise")
    ap.add_argument("--only", default=None, help="comma-separated substrings of kernel names to time")
    ap.add_argument("--pwc", action="store_true")
    ap.add_argument("--stock", action="store_true", help="also time the same ops through stock PyTorch CUDA kernels")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--all-precisions", action="store_true")
    ap.add_argument("--lib", default=None, help="path of an alternative build of libffcorr.so (kernel variants under test)")
    a = ap.parse_args()
    if a.lib:
        from focusflow_official_b200 import _lib as _L

        _L.LIB_PATH = os.path.abspath(a.lib)
    if a.all_precisions:
        for pr in ("fp16", "tf32", "bf16x3"):
            time_kernels(a.config, a.iters, pr, sigma=a.sigma, verbose=True, warmup=a.warmup, smooth=a.smooth,
                         only=a.only.split(',') if a.only else None)
    else:
        time_kernels(a.config, a.iters, a.precision, sigma=a.sigma, verbose=True, warmup=a.warmup, smooth=a.smooth,
                         only=a.only.split(',') if a.only else None)
    if a.stock:
        time_stock_aten(a.config, sigma=a.sigma, verbose=True)
    if a.pwc:
        time_pwc(a.iters, verbose=True, warmup=a.warmup)
