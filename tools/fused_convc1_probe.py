"""Development probe for the fused lookup + convc1 kernel (CorrBlock.lookup_conv): error against relu(conv1x1(lookup)) in
fp32, timing against lookup + cuDNN conv, and -- if the values are off -- one-hot weight diagnostics that show which sample
/ channel every operand slot is actually multiplied with."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import focusflow_official_b200 as ff  # noqa: E402

if os.environ.get("FFCORR_PROBE_LIB"):          # an alternative build of libffcorr.so (kernel variants under test)
    from focusflow_official_b200 import _lib as _L
    _L.LIB_PATH = os.path.abspath(os.environ["FFCORR_PROBE_LIB"])


def main():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = "cuda:0"
    shapes = [(1, 46, 62), (2, 24, 40), (8, 47, 156)] if len(sys.argv) < 2 else [tuple(int(x) for x in sys.argv[1].split("x"))]
    for (b, h, w) in shapes:
        torch.manual_seed(3)
        f1 = torch.randn(b, 256, h, w, device=dev) * 2.0
        f2 = torch.randn(b, 256, h, w, device=dev) * 2.0
        conv = torch.nn.Conv2d(324, 256, 1).to(dev)
        with torch.no_grad():
            conv.weight.mul_(2.0)
            conv.bias.uniform_(-0.5, 0.5)
        grid = ff.coords_grid(b, h, w, dev)
        flow = torch.nn.functional.interpolate(torch.randn(b, 2, h // 8 + 1, w // 8 + 1, device=dev) * 5, size=(h, w), mode="bilinear")
        coords = grid + flow
        coords[:, :, 0, 0] = -50.0          # a window entirely outside
        coords[:, :, -1, -1] += 0.5
        with torch.no_grad():
            blk = ff.CorrBlock(f1, f2, channels_last=True)
            v = blk(coords)                                      # [B, 324, h, w] channels_last
            ref = torch.relu(torch.nn.functional.conv2d(v, conv.weight, conv.bias))
            out = blk.lookup_conv(coords, conv)
            torch.cuda.synchronize()
            err = (out - ref).abs().max().item()
            scale = ref.abs().max().item()
            # yardstick: the TF32 convolution the reference runs
            torch.backends.cudnn.allow_tf32 = True
            ref_tf32 = torch.relu(torch.nn.functional.conv2d(v, conv.weight, conv.bias))
            torch.backends.cudnn.allow_tf32 = False
            err_tf32 = (ref_tf32 - ref).abs().max().item()
            print(json.dumps({"shape": [b, h, w], "max_abs_err": err, "max_ref": scale, "rel": err / scale,
                              "tf32_conv_max_abs_err": err_tf32, "finite": bool(torch.isfinite(out).all())}), flush=True)
            if err > 5e-3 * scale:
                # diagnostics with one-hot weights: out[:, co] should be relu(v[:, ci])
                vf = v.permute(0, 2, 3, 1).reshape(-1, 324)
                for (co, ci) in [(0, 0), (1, 1), (5, 9), (130, 10), (200, 81), (255, 323), (31, 100), (32, 101)]:
                    conv.weight.zero_(); conv.bias.zero_()
                    conv.weight[co, ci, 0, 0] = 1.0
                    o = blk.lookup_conv(coords, conv).permute(0, 2, 3, 1).reshape(-1, 256)
                    nz = (o.abs().sum(0) > 0).nonzero().flatten().tolist()
                    best = None
                    if nz:
                        col = o[:, nz[0]]
                        d = (torch.relu(vf) - col[:, None]).abs().mean(0)
                        best = (int(d.argmin()), float(d.min()))
                    print(json.dumps({"one_hot": [co, ci], "nonzero_out_channels": nz[:8], "n_nonzero": len(nz), "best_matching_sample": best}), flush=True)
                return
        if (b, h, w) == (8, 47, 156):
            def t(fn, n=20):
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
                ts = []
                for _ in range(n):
                    flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); fn(); e1.record(); e1.synchronize()
                    ts.append(e0.elapsed_time(e1) * 1e3)
                ts.sort()
                return round(ts[len(ts) // 2], 1), round(ts[0], 1)
            torch.backends.cudnn.allow_tf32 = True
            torch.backends.cudnn.benchmark = True
            conv_cl = conv.to(memory_format=torch.channels_last)
            with torch.no_grad():
                smooth = grid + flow                  # no artificial outliers: every unit takes the fast path
                print(json.dumps({"coords": "smooth flow only", "lookup_us": t(lambda: blk(smooth)),
                                  "lookup_plus_cudnn_conv_relu_us": t(lambda: torch.cudnn_convolution_relu(blk(smooth), conv_cl.weight, conv_cl.bias, (1, 1), (0, 0), (1, 1), 1)),
                                  "fused_us": t(lambda: blk.lookup_conv(smooth, conv))}), flush=True)
                print(json.dumps({"coords": "with 8 windows at -50 (integer coordinates -> exact per-tap path) and 8 at inf",
                                  "lookup_us": t(lambda: blk(coords)),
                                  "lookup_plus_cudnn_conv_relu_us": t(lambda: torch.cudnn_convolution_relu(blk(coords), conv_cl.weight, conv_cl.bias, (1, 1), (0, 0), (1, 1), 1)),
                                  "fused_us": t(lambda: blk.lookup_conv(coords, conv))}), flush=True)
                dump_trace()


def dump_trace():
    """Development builds with -DFFCORR_MO_TRACE: per-role clock stamps of the last launch, in us since the CTA started."""
    import ctypes

    import numpy as np
    from focusflow_official_b200 import _lib
    L = _lib.lib()
    if not hasattr(L, "ffcorr_debug_mo_trace"):
        return
    buf = np.zeros(148 * 128, dtype=np.int64)
    L.ffcorr_debug_mo_trace.argtypes = [ctypes.c_void_p]
    L.ffcorr_debug_mo_trace(buf.ctypes.data)
    tr = buf.reshape(148, 128)
    for cta in (0, 31, 32, 100, 147):
        t0 = tr[cta, 0]
        us = lambda s: round(float(tr[cta, s] - t0) / 1965.0, 2) if tr[cta, s] else None
        print(json.dumps({"cta": cta, "entry": us(120), "tmem_allocated": us(121), "smem_zeroed": us(122), "tmem_freed": us(123),
                          "weights_in_tmem": us(1), "mma_sees_weights": us(2), "end": us(127),
                          "mma[vfull,accempty,issued]": [[us(10 + i * 4), us(11 + i * 4), us(12 + i * 4)] for i in range(7)],
                          "gather_w0[start,done]": [[us(50 + i * 2), us(51 + i * 2)] for i in range(7)],
                          "gather_w7[start,done]": [[us(80 + i * 2), us(81 + i * 2)] for i in range(7)],
                          "epilogue[accfull,stored]": [[us(100 + i * 2), us(101 + i * 2)] for i in range(7)]}), flush=True)


if __name__ == "__main__":
    main()
